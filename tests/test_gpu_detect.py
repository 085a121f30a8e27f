"""GPU tier (run on the B200 box: ``pytest -m gpu``): the CUDA detect path, called through the
C ABI, against (i) the reference-generated golden fixtures, (ii) the numpy oracle on seeded
inputs, (iii) size-independent properties at BASELINE.json's full size.

Bars: kept-index sets / counts / masks bit-exact; decoded boxes within 1e-5 relative (fp32,
only ``exp`` differs between CUDA and the CPU libm — SURVEY.md A.1); pure copies exact.
"""
import numpy as np
import pytest
import torch

from oracle import box_oracle as bo
from tests import gen

pytestmark = pytest.mark.gpu

RTOL, ATOL = 1e-5, 1e-6          # north_star: decoded boxes within 1e-5 relative in fp32


def cu(a):
    return torch.as_tensor(np.ascontiguousarray(a)).cuda()


@pytest.fixture(scope='module')
def rd():
    import refinedet.pytorch_b200 as rd
    rd._ffi.lib()
    return rd


def load_detect(golden, tag):
    g = golden('detect_%s.npz' % tag)
    C, top_k, keep_top_k, conf_thr, nms_thr, obj_thr = g['params']
    return g, int(C), int(top_k), int(keep_top_k), float(conf_thr), float(nms_thr), float(obj_thr)


# ---------------------------------------------------------------------------------------------
# golden fixtures (outputs of the unmodified reference)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize('tag', ['sparse', 'dense'])
def test_forward_golden(rd, golden, tag):
    g, C, top_k, keep_top_k, conf_thr, nms_thr, obj_thr = load_detect(golden, tag)
    det = rd.Detect_RefineDet(C, 320, 0, top_k, conf_thr, nms_thr, obj_thr, keep_top_k)
    conf = cu(g['odm_conf'])
    boxes, scores = det.forward(cu(g['arm_loc']), cu(g['arm_conf']), cu(g['odm_loc']), conf, cu(g['priors']))
    np.testing.assert_allclose(boxes.cpu().numpy(), g['boxes'], rtol=RTOL, atol=ATOL)
    assert np.array_equal(scores.cpu().numpy(), g['scores'])
    assert np.array_equal(conf.cpu().numpy(), g['conf_after'])          # in-place ARM zeroing
    assert scores.data_ptr() != conf.data_ptr()


@pytest.mark.parametrize('tag', ['sparse', 'dense'])
def test_fused_detect_golden_a4(rd, golden, tag):
    g, C, top_k, keep_top_k, conf_thr, nms_thr, obj_thr = load_detect(golden, tag)
    det = rd.Detect_RefineDet(C, 320, 0, top_k, conf_thr, nms_thr, obj_thr, keep_top_k)
    conf = cu(g['odm_conf'])
    res = det.detect(cu(g['arm_loc']), cu(g['arm_conf']), cu(g['odm_loc']), conf, cu(g['priors']),
                     scale=g['scale'])
    assert np.array_equal(conf.cpu().numpy(), g['odm_conf'])            # fused stage never mutates inputs
    counts = res.counts.cpu().numpy()
    assert np.array_equal(counts, g['a4_counts'])
    dets = res.dets.cpu().numpy()
    B = counts.shape[0]
    for b in range(B):
        for c in range(C):
            n = counts[b, c]
            ref = g['a4_dets'][b, c, :n]
            assert np.array_equal(dets[b, c, :n, 4], ref[:, 4])          # scores are copies: exact
            np.testing.assert_allclose(dets[b, c, :n, :4], ref[:, :4], rtol=RTOL, atol=ATOL * 320)


@pytest.mark.parametrize('tag', ['sparse', 'dense'])
def test_forward_python_nms_golden_a5(rd, golden, tag):
    g, C, top_k, keep_top_k, conf_thr, nms_thr, obj_thr = load_detect(golden, tag)
    det = rd.Detect_RefineDet(C, 320, 0, top_k, conf_thr, nms_thr, obj_thr, keep_top_k)
    conf = cu(g['odm_conf'])
    out = det.forward_python_nms(cu(g['arm_loc']), cu(g['arm_conf']), cu(g['odm_loc']), conf, cu(g['priors']))
    out = out.cpu().numpy()
    ref = g['a5_output']
    assert out.shape == ref.shape
    assert np.array_equal(out[..., 0], ref[..., 0])                      # scores + zero padding exact
    np.testing.assert_allclose(out[..., 1:], ref[..., 1:], rtol=RTOL, atol=ATOL)
    assert np.array_equal(conf.cpu().numpy(), g['conf_after'])


def test_box_utils_golden(rd, golden):
    g = golden('box_utils.npz')
    bu = rd.box_utils
    pri, loc, truths, matched = cu(g['priors']), cu(g['loc']), cu(g['truths']), cu(g['matched'])
    assert np.array_equal(bu.point_form(pri).cpu().numpy(), g['point_form'])
    dec = bu.decode(loc, pri, [0.1, 0.2])
    np.testing.assert_allclose(dec.cpu().numpy(), g['decode'], rtol=RTOL, atol=ATOL)
    assert np.array_equal(bu.center_size(cu(g['decode'])).cpu().numpy(), g['center_size'])
    np.testing.assert_allclose(bu.encode(matched, pri, [0.1, 0.2]).cpu().numpy(), g['encode'], rtol=RTOL, atol=ATOL)
    pf = cu(g['point_form'])
    assert np.array_equal(bu.intersect(truths, pf).cpu().numpy(), g['intersect'])
    assert np.array_equal(bu.jaccard(truths, pf).cpu().numpy(), g['jaccard'])
    assert np.array_equal(bu.jaccard(truths, cu(g['decode'])).cpu().numpy(), g['jaccard_dec'])
    np.testing.assert_allclose(bu.log_sum_exp(cu(g['x'])).cpu().numpy(), g['log_sum_exp'], rtol=1e-6)
    # SURVEY.md Appendix B KAT: two-stage decode
    d1 = bu.decode(cu(g['kat_arm']), cu(g['kat_pri']), [0.1, 0.2])
    d2 = bu.decode(cu(g['kat_odm']), bu.center_size(d1), [0.1, 0.2])
    np.testing.assert_allclose(d2.cpu().numpy(), [[0.38644654, 0.45778579, 0.59460866, 0.55483037],
                                                  [-0.06584629, -0.00404091, 0.07186650, 0.07004091]], rtol=2e-6)


def test_nms_box_utils_golden(rd, golden):
    g = golden('nms_box_utils.npz')
    bu = rd.box_utils
    k, c = bu.nms(cu(g['kat_boxes']), cu(g['kat_scores']), 0.45, 200)
    assert c == 3 and k.tolist() == [3, 4, 5, 0, 0, 0]
    k, c = bu.nms(cu(g['kat_boxes']), cu(g['kat_scores']), 0.45, 3)
    assert c == 1 and k.tolist() == [3, 0, 0, 0, 0, 0]
    e = bu.nms(torch.zeros(0, 4).cuda(), torch.zeros(0).cuda(), 0.45, 200)
    assert isinstance(e, torch.Tensor) and e.numel() == 0 and e.dtype == torch.int64
    for tag in 'abc':
        thr, tk = g['args_' + tag]
        k, c = bu.nms(cu(g['boxes']), cu(g['scores']), float(thr), int(tk))
        assert c == int(g['count_' + tag])
        assert np.array_equal(k.cpu().numpy(), g['keep_' + tag])


def test_nms_wrapper_golden(rd, golden):
    g = golden('nms_pixel.npz')
    nms = rd.nms_wrapper.nms
    for tag, thr in (('045', 0.45), ('049', 0.49), ('070', 0.7)):
        assert nms(g['dets'], thr) == list(g['keep_' + tag])              # host ndarray -> rd_nms_host
        assert nms(cu(g['dets']), thr) == list(g['keep_' + tag])          # device tensor -> rd_nms
        assert nms(g['dets'], thr, force_cpu=True) == bo.nms_pixel(g['dets'], thr, suppress_on_equal=True)
    assert nms(np.zeros((0, 5), np.float32), 0.5) == []


# ---------------------------------------------------------------------------------------------
# oracle parity on seeded inputs
# ---------------------------------------------------------------------------------------------
def _oracle_a4(boxes, scores, scale, conf_thr, top_k, nms_thr, keep_top_k):
    B, P, C = scores.shape
    counts = np.zeros((B, C), np.int32)
    anchors = {}
    rows = {}
    for b in range(B):
        out, anc = bo.detect_stage_eval(boxes[b], scores[b], scale, conf_thr, top_k, nms_thr, keep_top_k)
        for c in range(C):
            counts[b, c] = out[c].shape[0]
            anchors[b, c] = anc[c]
            rows[b, c] = out[c]
    return counts, anchors, rows


@pytest.mark.parametrize('kind,B,size,C,top_k,keep', [
    ('sparse', 3, '320', 21, 1000, 500),
    ('dense', 2, '320', 21, 400, 200),       # top_k saturates: exercises the radix select + large kernel
    ('dense', 2, '320', 3, 400, 100),        # keep_top_k cap bites
    ('sparse', 2, '512', 81, 1000, 500),
    ('sparse-8', 2, '512', 2, 1000, 500),    # few classes: the wide (<= 1024 candidates, 256 threads) graph variant, n > 256
    ('sparse-8', 3, '512', 3, 1000, 120),    # ... with the keep_top_k cap biting
    ('sparse-8', 2, '512', 81, 1000, 500),   # the benchmark regime: ~700 nodes per image, common variant
    ('sparse-7', 2, '512', 81, 1000, 500),   # ~1900 nodes: two-block graph, > 256 candidates -> graph resolve in nms_large
    ('sparse-7', 2, '512', 3, 1000, 500),    # ... wide variant overflows 1024 candidates -> per-problem bins
    ('sparse-5', 2, '320', 21, 1000, 500),   # ~2700 nodes of 6375: three-block graph
    ('clustered', 2, '512', 81, 1000, 500),  # trained-detector-like: dozens of overlapping boxes per object
    ('clustered', 2, '320', 21, 1000, 500),
    ('clustered', 3, '512', 2, 1000, 500),
    ('clustered-tight', 2, '512', 21, 1000, 500),   # near-coincident boxes: degree > 64 -> no graph, bin path
])
def test_fused_detect_vs_oracle(rd, kind, B, size, C, top_k, keep):
    priors = rd.PriorBox(rd.REFINEDET_ANCHORS[size]).forward()
    P = priors.shape[0]
    arm_shift = {'sparse': -3.0, 'sparse-8': -8.0, 'sparse-7': -7.0, 'sparse-5': -5.0, 'dense': 0.0,
                 'clustered': 0.0, 'clustered-tight': 0.0}[kind]
    jitter, near_iou = (0.15, 0.2) if kind.endswith('tight') else (1.0, 0.35)
    kind = kind.split('-')[0]
    conf_thr, nms_thr, obj_thr = 0.01, 0.45, 0.01
    # fp32 softmax scores collide often at these candidate counts, so tie-free inputs are not
    # attainable by reseeding; the kernel and the oracle share one documented tie rule (score
    # descending, lower anchor first) and must agree bit-exactly WITH ties present.  The golden
    # fixtures (reference outputs) were generated tie-free, where the rule is unobservable.
    if kind == 'clustered':
        arm_loc, arm_conf, odm_loc, odm_conf = gen.detect_inputs_clustered(4242 + B + C, B, priors, C, jitter=jitter, near_iou=near_iou)
    else:
        arm_loc, arm_conf, odm_loc, odm_conf = gen.detect_inputs(4242 + B + C, B, P, C, kind, arm_shift=arm_shift)
    det = rd.Detect_RefineDet(C, int(size), 0, top_k, conf_thr, nms_thr, obj_thr, keep)
    scale = np.array([float(size)] * 4, np.float32)
    d_in = [t.cuda() for t in (arm_loc, arm_conf, odm_loc, odm_conf)]
    res = det.detect(*d_in, priors.cuda(), scale=scale)
    # a3 on the GPU gives the boxes the fused kernel used (same device function -> same bits)
    conf_a3 = odm_conf.clone().cuda()
    g_boxes, g_scores = det.forward(d_in[0], d_in[1], d_in[2], conf_a3, priors.cuda())
    # a3 vs oracle
    o_conf = odm_conf.numpy().copy()
    o_boxes, o_scores = bo.detect_forward(arm_loc.numpy(), arm_conf.numpy(), odm_loc.numpy(), o_conf,
                                          priors.numpy(), obj_thr)
    np.testing.assert_allclose(g_boxes.cpu().numpy(), o_boxes, rtol=RTOL, atol=ATOL)
    assert np.array_equal(g_scores.cpu().numpy(), o_scores)
    assert np.array_equal(conf_a3.cpu().numpy(), o_conf)
    # a4: oracle NMS fed the GPU's boxes -> kept anchor sets must be bit-exact
    counts, anchors, rows = _oracle_a4(g_boxes.cpu().numpy(), o_scores, scale, conf_thr, top_k, nms_thr, keep)
    g_counts = res.counts.cpu().numpy()
    assert np.array_equal(g_counts, counts)
    g_anchor = res.anchors.cpu().numpy()
    g_dets = res.dets.cpu().numpy()
    for b in range(B):
        for c in range(C):
            n = counts[b, c]
            assert np.array_equal(g_anchor[b, c, :n], anchors[b, c]), (b, c)
            assert np.array_equal(g_dets[b, c, :n], rows[b, c]), (b, c)
    assert (g_counts[:, 0] == 0).all()
    # and against the oracle's own boxes (exp differs by <= 2 ulp): same sets on these inputs
    counts2, anchors2, _ = _oracle_a4(o_boxes, o_scores, scale, conf_thr, top_k, nms_thr, keep)
    assert np.array_equal(g_counts, counts2)
    # second call on the same workspace gives identical results (workspace is left clean)
    res2 = det.detect(*d_in, priors.cuda(), scale=scale)
    assert torch.equal(res2.counts, res.counts)
    for b in range(B):
        for c in range(C):
            n = counts[b, c]
            assert torch.equal(res2.anchors[b, c, :n], res.anchors[b, c, :n])


@pytest.mark.parametrize('arm_shift,B,size,C', [(-8.0, 3, '320', 21), (-8.0, 2, '512', 81), (-6.0, 2, '320', 5)])
def test_fused_detect_logits_in(rd, arm_shift, B, size, C):
    """f-1: ``logits=True`` folds the softmax of models/refinedet.py:143-147 into the stage.
    Oracle = torch.softmax on the CPU followed by the a4 oracle.  The fused softmax reduces the row sum in
    a different order, so a score may differ from torch's in the last bits (tolerance 2e-6 relative,
    written here); a problem whose outcome could depend on those bits — a score within 1e-6 relative of
    the threshold or of another candidate's score, an ARM probability within 1e-6 relative of the
    objectness threshold — is excluded from the exact comparison, and at least 85 % must remain."""
    priors = rd.PriorBox(rd.REFINEDET_ANCHORS[size]).forward()
    P = priors.shape[0]
    conf_thr, nms_thr, obj_thr, top_k, keep = 0.01, 0.45, 0.01, 1000, 500
    arm_loc, arm_lg, odm_loc, odm_lg = gen.detect_logits(991 + C, B, P, C, 'sparse', arm_shift=arm_shift)
    arm_conf, odm_conf = torch.softmax(arm_lg, -1), torch.softmax(odm_lg, -1)
    det = rd.Detect_RefineDet(C, int(size), 0, top_k, conf_thr, nms_thr, obj_thr, keep)
    scale = np.array([float(size)] * 4, np.float32)
    res = det.detect(arm_loc.cuda(), arm_lg.cuda(), odm_loc.cuda(), odm_lg.cuda(), priors.cuda(), scale=scale,
                     logits=True)
    # reference: probabilities from torch.softmax, boxes from the GPU a3 kernel (same device function)
    conf_a3 = odm_conf.clone().cuda()
    g_boxes, g_scores = det.forward(arm_loc.cuda(), arm_conf.cuda(), odm_loc.cuda(), conf_a3, priors.cuda())
    o_scores = g_scores.cpu().numpy()
    counts, anchors, rows = _oracle_a4(g_boxes.cpu().numpy(), o_scores, scale, conf_thr, top_k, nms_thr, keep)
    g_counts = res.counts.cpu().numpy()
    g_anchor = res.anchors.cpu().numpy()
    g_dets = res.dets.cpu().numpy()
    eps = 1e-6
    p1 = torch.softmax(arm_lg.double(), -1)[..., 1].numpy()
    s64 = torch.softmax(odm_lg.double(), -1).numpy()
    checked = total = 0
    for b in range(B):
        arm_fragile = bool((np.abs(p1[b] - obj_thr) <= eps * obj_thr).any())
        passing = p1[b] > obj_thr
        for c in range(1, C):
            total += 1
            v = np.sort(s64[b, passing, c])
            v = v[v > conf_thr * (1 - 2 * eps)]
            fragile = arm_fragile or bool((np.abs(v - conf_thr) <= eps * conf_thr).any()) or \
                bool((np.diff(v) <= eps * v[1:]).any())
            if fragile:
                continue
            checked += 1
            n = counts[b, c]
            assert g_counts[b, c] == n, (b, c)
            assert np.array_equal(g_anchor[b, c, :n], anchors[b, c]), (b, c)
            assert np.array_equal(g_dets[b, c, :n, :4], rows[b, c][:, :4]), (b, c)
            np.testing.assert_allclose(g_dets[b, c, :n, 4], rows[b, c][:, 4], rtol=2e-6, atol=0)
    assert checked >= 0.85 * total, (checked, total)
    assert (g_counts[:, 0] == 0).all()


def _same_detections(a, b):
    assert torch.equal(a.counts, b.counts)
    m = torch.arange(a.dets.shape[2], device=a.dets.device).view(1, 1, -1) < a.counts.unsqueeze(-1)
    assert torch.equal(a.dets[m], b.dets[m])
    assert torch.equal(a.anchors[m], b.anchors[m])


def test_plan_replay_and_lanes(rd):
    """rd_detect_plan_*: the captured launch chain gives the results of the direct call, replay after replay,
    on another stream, with refilled input buffers, and with two plans in flight over separate lanes."""
    size, C, B = '320', 21, 4
    priors = rd.PriorBox(rd.REFINEDET_ANCHORS[size]).forward().cuda()
    P = priors.shape[0]
    det = rd.Detect_RefineDet(C, 320, 0, 1000, 0.01, 0.45, 0.01, 500)
    scale = torch.tensor([320.0] * 4).cuda()
    sets = [[t.cuda() for t in gen.detect_inputs(500 + i, B, P, C, 'sparse', arm_shift=-4.0)] for i in range(3)]
    direct = [det.detect(*a, priors, scale=scale) for a in sets]
    buf = [t.clone() for t in sets[0]]                               # static input buffers of the plan
    plan = det.plan(*buf, priors, scale=scale)
    _same_detections(plan.launch(), direct[0])
    side = torch.cuda.Stream()
    for i in (1, 2, 0):
        for d, s_ in zip(buf, sets[i]):
            d.copy_(s_)
        side.wait_stream(torch.cuda.current_stream())
        res = plan.launch(side)
        side.synchronize()
        _same_detections(res, direct[i])
    # two lanes, two plans, both in flight
    lanes = [(torch.cuda.Stream(), det.new_workspace(B, P, priors.device), det.new_outputs(B, priors.device))
             for _ in range(2)]
    plans = [det.plan(*sets[i], priors, scale=scale, workspace=lanes[i][1], out=lanes[i][2]) for i in range(2)]
    torch.cuda.synchronize()
    for rep in range(3):
        out = [plans[i].launch(lanes[i][0]) for i in range(2)]
    torch.cuda.synchronize()
    for i in range(2):
        _same_detections(out[i], direct[i])
    with pytest.raises(ValueError):
        det.plan(sets[0][0][:, ::2], sets[0][1], sets[0][2], sets[0][3], priors)      # non-contiguous: referenced, not copied
    plan.close()
    with pytest.raises(RuntimeError):
        plan.launch()


@pytest.mark.parametrize('dma_rows,B,C', [(True, 3, 21), (False, 3, 21), (True, 4, 24), (True, 32, 81)])
def test_host_pipeline(rd, dma_rows, B, C):
    """DetectHostPipeline (pinned host inputs, several lanes) == detect() + packed() on device copies.
    (4, 24) and (32, 81) meet the conditions of the line-granular PCIe row fetch of collect_kernel
    (B*P*C a multiple of 32), with row offsets of every residue inside the 128-byte lines."""
    size = '320'
    priors = rd.PriorBox(rd.REFINEDET_ANCHORS[size]).forward().cuda()
    P = priors.shape[0]
    det = rd.Detect_RefineDet(C, 320, 0, 1000, 0.01, 0.45, 0.01, 500)
    scale = torch.tensor([320.0] * 4).cuda()
    hs = [[t.pin_memory() for t in gen.detect_inputs(700 + i, B, P, C, 'sparse', arm_shift=-4.0)] for i in range(3)]
    pipe = rd.DetectHostPipeline(det, priors, scale, B, lanes=2, dma_rows=dma_rows)
    tickets = []
    got = []
    for k in range(6):
        tickets.append((k % 3, pipe.submit(hs[k % 3])))
        if len(tickets) == 2:
            i, t = tickets.pop(0)
            c, r = pipe.result(t)
            got.append((i, c.clone(), r.clone()))
    while tickets:
        i, t = tickets.pop(0)
        c, r = pipe.result(t)
        got.append((i, c.clone(), r.clone()))
    for i, c, r in got:
        ref = det.detect(*[t.cuda() for t in hs[i]], priors, scale=scale)
        offs, rows = ref.packed()
        assert torch.equal(c, ref.counts.cpu())
        assert torch.equal(r, rows.cpu())
    c2, r2 = det.detect_host(hs[0], priors, scale)                   # serial form
    assert torch.equal(c2, got[0][1]) and torch.equal(r2, got[0][2])
    # one lane, batches of growing and shrinking row counts through the same lane buffers
    if dma_rows:
        dense = [t.pin_memory() for t in gen.detect_inputs(990, B, P, C, 'sparse', arm_shift=-2.0)]
        pipe1 = rd.DetectHostPipeline(det, priors, scale, B, lanes=1)
        for h in (hs[0], dense, hs[1], dense, dense, hs[2]):
            c, r = pipe1.result(pipe1.submit(h))
            ref = det.detect(*[t.cuda() for t in h], priors, scale=scale)
            assert torch.equal(c, ref.counts.cpu()) and torch.equal(r, ref.packed()[1].cpu())


def test_forward_python_nms_vs_oracle(rd):
    size, C, B, top_k = '320', 5, 2, 200
    priors = rd.PriorBox(rd.REFINEDET_ANCHORS[size]).forward()
    P = priors.shape[0]
    arm_loc, arm_conf, odm_loc, odm_conf = gen.detect_inputs(77, B, P, C, 'dense')
    det = rd.Detect_RefineDet(C, 320, 0, top_k, 0.01, 0.45, 0.01, 50)
    conf = odm_conf.clone().cuda()
    out = det.forward_python_nms(arm_loc.cuda(), arm_conf.cuda(), odm_loc.cuda(), conf, priors.cuda())
    conf2 = odm_conf.clone().cuda()
    g_boxes, g_scores = det.forward(arm_loc.cuda(), arm_conf.cuda(), odm_loc.cuda(), conf2, priors.cuda())
    gb, gs = g_boxes.cpu().numpy(), g_scores.cpu().numpy()
    exp = np.zeros((B, C, top_k, 5), np.float32)
    for i in range(B):
        for cl in range(1, C):
            m = gs[i, :, cl] > np.float32(0.01)
            if m.sum() == 0:
                continue
            ids, count = bo.nms(gb[i][m], gs[i, m, cl], 0.45, top_k)
            ids = ids[:count]
            exp[i, cl, :count, 0] = gs[i, m, cl][ids]
            exp[i, cl, :count, 1:] = gb[i][m][ids]
    assert np.array_equal(out.cpu().numpy(), exp)
    assert torch.equal(conf, conf2)


@pytest.mark.parametrize('n,top_k,flavour', [(1, 5, 0), (33, 200, 0), (700, 200, 1), (3000, 1000, 1),
                                             (4096, 4096, 0), (5000, 300, 1)])
def test_standalone_nms_vs_oracle(rd, n, top_k, flavour):
    g = torch.Generator().manual_seed(n)
    if flavour:
        cxy = 512 * torch.rand(n, 2, generator=g)
        wh = 10 + 90 * torch.rand(n, 2, generator=g)
    else:
        cxy = torch.rand(n, 2, generator=g)
        wh = 0.03 + 0.2 * torch.rand(n, 2, generator=g)
    boxes = torch.cat([cxy - wh / 2, cxy + wh / 2], 1)
    scores = torch.rand(n, generator=g)
    assert scores.unique().numel() == n
    if flavour:
        dets = torch.cat([boxes, scores[:, None]], 1).numpy()
        order = np.argsort(-dets[:, 4], kind='stable')[:top_k]
        exp = [int(order[i]) for i in bo.nms_pixel(dets[order], 0.45)]
        from refinedet.pytorch_b200 import _ffi
        keep, count = rd.box_utils.nms_device(boxes.cuda(), scores.cuda(), 0.45, top_k, _ffi.RD_NMS_PIXEL_PLUS1)
        assert keep[:int(count)].tolist() == exp
    else:
        ek, ec = bo.nms(boxes.numpy(), scores.numpy(), 0.45, top_k)
        keep, count = rd.box_utils.nms(boxes.cuda(), scores.cuda(), 0.45, top_k)
        assert count == ec
        assert np.array_equal(keep.cpu().numpy(), ek)


def test_nms_degenerate_boxes_and_thresholds(rd):
    """zero-area / inverted / duplicate boxes and thr = 0 take the exact path (no spatial cull)."""
    boxes = np.array([[0.1, 0.1, 0.5, 0.5], [0.1, 0.1, 0.5, 0.5], [0.7, 0.7, 0.7, 0.9], [0.2, 0.2, 0.2, 0.2],
                      [0.9, 0.9, 0.8, 0.8], [0.6, 0.1, 0.9, 0.4], [0.61, 0.11, 0.9, 0.4], [0.3, 0.3, 0.6, 0.6]],
                     np.float32)
    scores = np.array([0.9, 0.8, 0.7, 0.6, 0.5, 0.4, 0.3, 0.2], np.float32)
    for thr in (0.0, 0.3, 0.45, 0.99):
        ek, ec = bo.nms(boxes, scores, thr, 200)
        keep, count = rd.box_utils.nms(cu(boxes), cu(scores), thr, 200)
        assert count == ec and np.array_equal(keep.cpu().numpy(), ek), thr
    dets = np.concatenate([boxes * 300, scores[:, None]], 1).astype(np.float32)
    for thr in (0.0, 0.3, 0.7):
        assert rd.nms_wrapper.nms(dets, thr) == bo.nms_pixel(dets, thr)
        assert rd.nms_wrapper.nms(dets, thr, force_cpu=True) == bo.nms_pixel(dets, thr, suppress_on_equal=True)


def test_nms_heavy_overlap_clusters(rd):
    """many near-duplicates: long suppression chains, kept << n (the cull gives no help)."""
    g = torch.Generator().manual_seed(5)
    centres = torch.rand(12, 2, generator=g) * 400 + 50
    idx = torch.randint(0, 12, (2000,), generator=g)
    cxy = centres[idx] + 6 * torch.randn(2000, 2, generator=g)
    wh = 60 + 10 * torch.rand(2000, 2, generator=g)
    dets = torch.cat([cxy - wh / 2, cxy + wh / 2, torch.rand(2000, 1, generator=g)], 1).numpy()
    for thr in (0.3, 0.49, 0.8):
        assert rd.nms_wrapper.nms(dets, thr) == bo.nms_pixel(dets, thr)


# ---------------------------------------------------------------------------------------------
# full size (BASELINE.json config 3): properties that do not need the oracle
# ---------------------------------------------------------------------------------------------
def test_full_size_properties(rd):
    B, C, top_k, keep = 32, 81, 1000, 500
    priors = rd.PriorBox(rd.REFINEDET_ANCHORS['512']).forward().cuda()
    P = priors.shape[0]
    arm_loc, arm_conf, odm_loc, odm_conf = [t.cuda() for t in gen.detect_inputs(1234 + 3000, B, P, C, 'sparse')]
    det = rd.Detect_RefineDet(C, 512, 0, top_k, 0.01, 0.45, 0.01, keep)
    scale = torch.tensor([512.0] * 4)
    res = det.detect(arm_loc, arm_conf, odm_loc, odm_conf, priors, scale=scale)
    counts = res.counts
    assert int(counts[:, 0].sum()) == 0 and int(counts.max()) <= keep and int(counts.sum()) > 0
    conf_copy = odm_conf.clone()
    boxes, scores = det.forward(arm_loc, arm_conf, odm_loc, conf_copy, priors)
    # every emitted row is (scaled box of its anchor, score of its anchor/class), scores descending
    ar = torch.arange(keep, device='cuda')
    valid = ar[None, None, :] < counts[:, :, None]
    a = res.anchors.long().clamp(min=0, max=P - 1)
    bidx = torch.arange(B, device='cuda')[:, None, None].expand_as(a)
    cidx = torch.arange(C, device='cuda')[None, :, None].expand_as(a)
    exp_score = scores[bidx, a, cidx]
    exp_box = boxes[bidx, a] * scale.cuda()
    assert torch.equal(res.dets[..., 4][valid], exp_score[valid])
    assert torch.equal(res.dets[..., :4][valid], exp_box[valid])
    assert bool((res.dets[..., 4][valid] > 0.01).all())
    s = res.dets[..., 4]
    desc = (s[..., 1:] <= s[..., :-1]) | ~valid[..., 1:]
    assert bool(desc.all())
    # NMS invariant: no two kept rows of one (image, class) overlap above the threshold (+1 IoU)
    for b in (0, 17, 31):
        for c in (1, 40, 80):
            n = int(counts[b, c])
            d = res.dets[b, c, :n, :4]
            x1, y1, x2, y2 = d[:, 0], d[:, 1], d[:, 2], d[:, 3]
            area = (x2 - x1 + 1) * (y2 - y1 + 1)
            w = (torch.min(x2[:, None], x2[None]) - torch.max(x1[:, None], x1[None]) + 1).clamp(min=0)
            h = (torch.min(y2[:, None], y2[None]) - torch.max(y1[:, None], y1[None]) + 1).clamp(min=0)
            iou = w * h / (area[:, None] + area[None] - w * h)
            iou.fill_diagonal_(0)
            assert float(iou.max()) <= 0.45 if n else True
    # idempotence / determinism and pack round trip
    res2 = det.detect(arm_loc, arm_conf, odm_loc, odm_conf, priors, scale=scale)
    assert torch.equal(res2.counts, counts)
    assert torch.equal(res2.dets[valid], res.dets[valid])
    offsets, rows = res.packed()
    assert int(offsets[-1]) == int(counts.sum()) == rows.shape[0]
    assert torch.equal(rows, res.dets[valid])
    assert torch.equal(offsets[:-1].long(), (counts.flatten().long().cumsum(0) - counts.flatten().long()))
    # one oracle spot check at full size: image 5, three classes
    gb, gs = boxes[5].cpu().numpy(), scores[5].cpu().numpy()
    out, anc = bo.detect_stage_eval(gb, gs, scale.numpy(), 0.01, top_k, 0.45, keep)
    for c in (1, 33, 80):
        n = int(counts[5, c])
        assert np.array_equal(res.anchors[5, c, :n].cpu().numpy(), anc[c])


@pytest.mark.parametrize('kind,B,P,C', [('sparse', 32, 16320, 81), ('dense', 32, 16320, 81), ('sparse', 5, 6375, 21),
                                         ('sparse', 3, 1000, 2), ('dense', 7, 777, 33)])
def test_forward_full_size(rd, kind, B, P, C):
    """a3 at BASELINE.json's sizes: scores / in-place zeroing bit-exact against the masked copy the reference's
    ``odm_conf_data[no_object_index] = 0`` amounts to (detection_refinedet.py:40-42), the scalar path of the kernel
    (a misaligned view of the same data) against the 16-byte path, row counts that are no multiple of 32, and a
    second call on the already zeroed tensor (idempotent)."""
    arm_loc, arm_conf, odm_loc, odm_conf = [t.cuda() for t in gen.detect_inputs(99 + B + C, B, P, C, kind)]
    priors = torch.rand(P, 4).cuda() * 0.5 + 0.05
    det = rd.Detect_RefineDet(C, 512, 0, 1000, 0.01, 0.45, 0.01, 500)
    conf = odm_conf.clone()
    boxes, scores = det.forward(arm_loc, arm_conf, odm_loc, conf, priors)
    keep = ~(arm_conf[..., 1] <= 0.01)
    want = odm_conf * keep.unsqueeze(-1)
    assert torch.equal(scores, want)
    assert torch.equal(conf, want)
    frac = float(keep.float().mean())
    assert (frac < 0.2) if kind == 'sparse' else (frac > 0.5)
    # a misaligned copy of the same data (4-byte offset: no 16-byte alignment, the kernel's scalar path)
    pad = torch.empty(odm_conf.numel() + 1, device='cuda')
    conf2 = pad[1:].view_as(odm_conf)
    conf2.copy_(odm_conf)
    pad_s = torch.empty(odm_conf.numel() + 1, device='cuda')
    scores2 = pad_s[1:].view_as(odm_conf)
    boxes2 = torch.empty_like(boxes)
    ffi = rd._ffi
    ffi.check(ffi.lib().rd_detect_forward(ffi.ptr(arm_loc), ffi.ptr(arm_conf), ffi.ptr(odm_loc), ffi.ptr(conf2), ffi.ptr(priors),
                                          B, P, C, 0.01, det.variance[0], det.variance[1], ffi.ptr(boxes2), ffi.ptr(scores2),
                                          ffi.stream_ptr()), 'rd_detect_forward')
    assert torch.equal(boxes2, boxes) and torch.equal(scores2, want) and torch.equal(conf2, want)
    boxes3, scores3 = det.forward(arm_loc, arm_conf, odm_loc, conf, priors)
    assert torch.equal(scores3, want) and torch.equal(conf, want) and torch.equal(boxes3, boxes)


def test_detect_host_zero_copy_matches_staged(rd):
    """e2e API: kernels reading the pinned host tensors directly give the same packed rows as the
    staged H2D copy path."""
    B, C = 3, 21
    priors = rd.PriorBox(rd.REFINEDET_ANCHORS['320']).forward().cuda()
    P = priors.shape[0]
    host = [t.pin_memory() for t in gen.detect_inputs(31, B, P, C, 'sparse', arm_shift=-3.0)]
    det = rd.Detect_RefineDet(C, 320, 0, 1000, 0.01, 0.45, 0.01, 500)
    scale = torch.tensor([320.0] * 4)
    c1, r1 = det.detect_host(host, priors, scale, zero_copy=False)
    c2, r2 = det.detect_host(host, priors, scale, zero_copy=True)
    assert torch.equal(c1, c2) and torch.equal(r1, r2) and int(c1.sum()) > 0
    with pytest.raises(RuntimeError):
        det.detect_host([t.clone() for t in host], priors, scale, zero_copy=True)     # not pinned


# ---------------------------------------------------------------------------------------------
# edge cases: ragged / tiny / empty inputs, extreme parameters
# ---------------------------------------------------------------------------------------------
def _run_vs_oracle(rd, B, P, C, top_k, keep, kind='dense', seed=5, conf_thr=0.01, nms_thr=0.45, obj_thr=0.01,
                   arm_shift=-3.0, size=320):
    g = torch.Generator().manual_seed(seed)
    cxy = torch.rand(P, 2, generator=g)
    wh = 0.03 + 0.3 * torch.rand(P, 2, generator=g)
    priors = torch.cat([cxy, wh], 1).clamp(0, 1)
    arm_loc, arm_conf, odm_loc, odm_conf = gen.detect_inputs(seed, B, P, C, kind, arm_shift=arm_shift)
    det = rd.Detect_RefineDet(C, size, 0, top_k, conf_thr, nms_thr, obj_thr, keep)
    scale = np.array([float(size)] * 4, np.float32)
    d_in = [t.cuda() for t in (arm_loc, arm_conf, odm_loc, odm_conf)]
    res = det.detect(*d_in, priors.cuda(), scale=scale)
    g_boxes, g_scores = det.forward(d_in[0], d_in[1], d_in[2], odm_conf.clone().cuda(), priors.cuda())
    gb, gs = g_boxes.cpu().numpy(), g_scores.cpu().numpy()
    counts = res.counts.cpu().numpy()
    anchors = res.anchors.cpu().numpy()
    dets = res.dets.cpu().numpy()
    for b in range(B):
        out, anc = bo.detect_stage_eval(gb[b], gs[b], scale, conf_thr, top_k, nms_thr, keep)
        for c in range(C):
            assert counts[b, c] == out[c].shape[0], (b, c, counts[b, c], out[c].shape[0])
            assert np.array_equal(anchors[b, c, :counts[b, c]], anc[c]), (b, c)
            assert np.array_equal(dets[b, c, :counts[b, c]], out[c]), (b, c)
    return counts


@pytest.mark.parametrize('B,P,C,top_k,keep', [
    (1, 1, 2, 10, 10),          # a single anchor
    (1, 33, 3, 1000, 500),      # one warp + one lane
    (2, 100, 5, 5, 5),          # top_k < candidates: radix-select path on a graph image
    (3, 1025, 4, 200, 1),       # slice boundary (1024 + 1), keep_top_k = 1
    (2, 2500, 2, 300, 300),     # 2-class (SAR config), >256 candidates per problem, 3 slices
    (1, 700, 128, 50, 20),      # the class limit (C = 128)
])
def test_edge_shapes_vs_oracle(rd, B, P, C, top_k, keep):
    counts = _run_vs_oracle(rd, B, P, C, top_k, keep)
    assert counts[:, 0].sum() == 0


def test_no_candidates_and_all_candidates(rd):
    # every anchor ARM-filtered: all counts zero, nothing else touched
    counts = _run_vs_oracle(rd, 2, 500, 4, 100, 50, obj_thr=2.0)
    assert counts.sum() == 0
    # conf threshold above every score
    counts = _run_vs_oracle(rd, 2, 500, 4, 100, 50, conf_thr=1.5)
    assert counts.sum() == 0
    # negative thresholds: every (anchor, class>=1) is a candidate
    counts = _run_vs_oracle(rd, 1, 300, 3, 1000, 500, conf_thr=-1.0, obj_thr=-1.0)
    assert counts[:, 1:].sum() > 0


def test_graph_fallback_paths_agree(rd):
    """An image with more nodes than the graph holds (4096), or whose graph overflows a node's 8 adjacency
    slots, falls back to per-problem bins; every path must give the oracle's result."""
    _run_vs_oracle(rd, 2, 6000, 3, 150, 100, kind='dense')                      # ~5200 nodes: no graph
    _run_vs_oracle(rd, 2, 3000, 3, 150, 100, kind='dense')                      # ~2600 nodes, 3 graph blocks, dense overlaps
    _run_vs_oracle(rd, 2, 3000, 3, 150, 100, kind='sparse', arm_shift=-2.0)     # few hundred nodes: graph


def test_limits_raise(rd):
    det = rd.Detect_RefineDet(200, 320, 0, 10, 0.01, 0.45, 0.01, 5)
    P = 64
    z = lambda *s: torch.zeros(*s).cuda()
    with pytest.raises(RuntimeError):
        det.detect(z(1, P, 4), z(1, P, 2), z(1, P, 4), z(1, P, 200), z(P, 4))    # C > 128
    det = rd.Detect_RefineDet(3, 320, 0, 5000, 0.01, 0.45, 0.01, 5)
    with pytest.raises(RuntimeError):
        det.detect(z(1, P, 4), z(1, P, 2), z(1, P, 4), z(1, P, 3), z(P, 4))      # top_k > 4096


def test_model_cfg1_fixture_gpu(rd, golden):
    """BASELINE.json config 1 (real RefineDet320/VOC head outputs, random init): a3 against the reference,
    a4 against the reference for tie-free classes and against the oracle for all (4,600 candidates per
    class: radix-select + large-problem path on real model data)."""
    g = golden('model_cfg1.npz')
    C, top_k, keep, conf_thr, nms_thr, obj_thr = [float(v) for v in g['params']]
    C, top_k, keep = int(C), int(top_k), int(keep)
    det = rd.Detect_RefineDet(C, 320, 0, top_k, conf_thr, nms_thr, obj_thr, keep)
    ins = [cu(g[k][None]) for k in ('arm_loc', 'arm_conf', 'odm_loc', 'odm_conf')]
    pri = cu(g['priors'])
    conf = ins[3].clone()
    boxes, scores = det.forward(ins[0], ins[1], ins[2], conf, pri)
    np.testing.assert_allclose(boxes[0].cpu().numpy(), g['boxes'], rtol=RTOL, atol=ATOL)
    assert np.array_equal(scores[0].cpu().numpy(), g['scores'])
    scale = np.array([320.0] * 4, np.float32)
    res = det.detect(*ins, pri, scale=scale)
    counts = res.counts.cpu().numpy()[0]
    dets = res.dets.cpu().numpy()[0]
    out, anc = bo.detect_stage_eval(boxes[0].cpu().numpy(), g['scores'], scale, conf_thr, top_k, nms_thr, keep)
    for j in range(1, C):
        assert counts[j] == out[j].shape[0], j
        assert np.array_equal(res.anchors[0, j, :counts[j]].cpu().numpy(), anc[j]), j
        if g['a4_tie_free'][j]:
            n = int(g['a4_counts'][j])
            assert counts[j] == n
            assert np.array_equal(dets[j, :n, 4], g['a4_dets'][j, :n, 4])
            np.testing.assert_allclose(dets[j, :n, :4], g['a4_dets'][j, :n, :4], rtol=RTOL, atol=ATOL * 320)


def test_api_variants(rd):
    """per-image scale, non-contiguous / misaligned inputs (copied by the wrapper, in-place contract kept),
    to_all_boxes layout."""
    B, C = 2, 4
    priors = rd.PriorBox(rd.REFINEDET_ANCHORS['320']).forward()[::3].contiguous().cuda()
    P = priors.shape[0]
    arm_loc, arm_conf, odm_loc, odm_conf = [t.cuda() for t in gen.detect_inputs(3, B, P, C, 'sparse', arm_shift=-2.0)]
    det = rd.Detect_RefineDet(C, 320, 0, 200, 0.01, 0.45, 0.01, 100)
    scale = torch.tensor([[320., 320., 320., 320.], [500., 375., 500., 375.]])
    res = det.detect(arm_loc, arm_conf, odm_loc, odm_conf, priors, scale=scale)
    boxes, scores = det.forward(arm_loc, arm_conf, odm_loc, odm_conf.clone(), priors)
    for b in range(B):
        out, _ = bo.detect_stage_eval(boxes[b].cpu().numpy(), scores[b].cpu().numpy(), scale[b].numpy(), 0.01, 200,
                                      0.45, 100)
        ab = res.to_all_boxes()
        for c in range(C):
            assert np.array_equal(ab[c][b], out[c].reshape(-1, 5)), (b, c)
    # non-contiguous odm_conf view + misaligned arm_loc: results equal, caller's tensor still zeroed in place
    big = torch.zeros(B, P, C + 3, device='cuda')
    big[..., :C] = odm_conf
    view = big[..., :C]
    assert not view.is_contiguous()
    raw = torch.zeros(arm_loc.numel() + 1, device='cuda')
    mis = raw[1:].view_as(arm_loc)
    mis.copy_(arm_loc)
    assert mis.data_ptr() % 16 != 0
    b2, s2 = det.forward(mis, arm_conf, odm_loc, view, priors)
    assert torch.equal(b2, boxes) and torch.equal(s2, scores)
    assert torch.equal(view, scores)                              # in-place zeroing reached the caller's storage


def test_coco_wire_format(rd):
    """SURVEY f-2: result rows in the reference's COCO json layout (data/sarship_coco.py:293-336)."""
    B, C = 3, 5
    priors = rd.PriorBox(rd.REFINEDET_ANCHORS['320']).forward()[::7].contiguous().cuda()
    P = priors.shape[0]
    ins = [t.cuda() for t in gen.detect_inputs(12, B, P, C, 'sparse', arm_shift=-2.0)]
    det = rd.Detect_RefineDet(C, 320, 0, 100, 0.01, 0.49, 0.01, 50)
    res = det.detect(*ins, priors, scale=[320., 320., 320., 320.])
    ids = [101, 7, 55]
    cats = [None, 3, 1, 18, 44]
    got = res.to_coco_results(ids, cats)
    exp = bo.coco_results(res.to_all_boxes(), ids, cats)
    assert len(got) == len(exp) == int(res.counts.sum()) > 0
    for a, b in zip(got, exp):
        assert a['image_id'] == b['image_id'] and a['category_id'] == b['category_id']
        assert a['bbox'] == b['bbox'] and a['score'] == b['score']
    import json
    json.dumps(got)                                              # serialisable like the reference's file


def test_coco_records_device_full_scale(rd):
    """f-2 at config-3 scale (B = 32, C = 81: ~390 k rows): the device-built records (rd_coco_records) equal the
    reference's host loop (oracle ``coco_results`` = data/sarship_coco.py:293-336), from the slot layout and from
    the packed (gathered) rows; classes mapped to None are skipped."""
    import time
    from refinedet.pytorch_b200 import dist as rdist
    B, C = 32, 81
    priors = rd.PriorBox(rd.REFINEDET_ANCHORS['512']).forward().cuda()
    ins = [t.cuda() for t in gen.detect_inputs(4234, B, priors.shape[0], C, 'sparse')]
    det = rd.Detect_RefineDet(C, 512, 0, 1000, 0.01, 0.45, 0.01, 500)
    res = det.detect(*ins, priors, scale=[512.] * 4)
    ids = [1000 + 3 * b for b in range(B)]
    cats = [None] + [c + 10 for c in range(1, C)]
    cats[7] = None                                                # an unmapped class
    t0 = time.perf_counter()
    got = res.to_coco_results(ids, cats)
    t_dev = time.perf_counter() - t0
    ab = res.to_all_boxes()
    ab[7] = [np.empty((0, 5), np.float32) for _ in range(B)]
    t0 = time.perf_counter()
    exp = bo.coco_results(ab, ids, cats)
    t_host = time.perf_counter() - t0
    assert len(got) == len(exp) > 300000
    assert got == exp                                             # ids, categories, float64 boxes, scores: all equal
    print('coco records: device path %.3f s, host double loop %.3f s, %d rows' % (t_dev, t_host, len(got)))
    # arrays only (no dicts): the form a json / pycocotools writer consumes
    a_ids, a_vals = res.to_coco_arrays(cats)
    assert a_ids.shape == (len(exp), 2) and a_vals.dtype == np.float64
    t0 = time.perf_counter()
    arr7 = res.to_coco_numpy(ids, cats)                           # pycocotools loadRes(ndarray) layout
    t_np = time.perf_counter() - t0
    assert arr7.shape == (len(exp), 7)
    assert np.array_equal(arr7[:, 0], [e['image_id'] for e in exp]) and np.array_equal(arr7[:, 6], [e['category_id'] for e in exp])
    assert np.array_equal(arr7[:, 1:5], np.array([e['bbox'] for e in exp])) and np.array_equal(arr7[:, 5], [e['score'] for e in exp])
    print('coco records as one [n,7] array: %.3f s' % t_np)
    # the same from packed rows (what the multi-GPU gather delivers), split as two "ranks"
    offs, rows = res.packed()
    offs = offs.long()
    half = B // 2
    cut = int(offs[half * C])
    g_ids, g_vals = rdist.gathered_to_coco_arrays([res.counts[:half], res.counts[half:]], [rows[:cut], rows[cut:]], cats)
    assert np.array_equal(g_ids, a_ids) and np.array_equal(g_vals, a_vals)
