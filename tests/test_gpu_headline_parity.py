"""GPU tier: oracle parity on the configurations the metric is quoted on, EVERY (image, class) problem.

The headline (BASELINE.json config 3: B = 32, P = 16,320, C = 81 — ``bench.py``'s exact input) runs the
``nms_small_kernel<256, 128>`` instantiation, config 5 (C = 2) the ``<1024, 256>`` one, trained-detector-like
(clustered) inputs the triangular bit-row resolve, the dense generator the queued large-problem kernel.  Each
is compared with ``bo.detect_stage_eval`` (eval_refinedet_coco.py:213-232 + utils/nms/py_cpu_nms.py:10-38) on
all B x C problems — counts, anchor lists and rows bit-exact — with the images fanned out over a process pool
(``tests/oracle_pool.py``).  A test-only flag (``instance=``) forces either instantiation so that both are also
held to the oracle where the heuristic would not pick them.
"""
import numpy as np
import pytest
import torch

from oracle import box_oracle as bo
from tests import gen
from tests.oracle_pool import assert_detections_equal, detect_stage_eval_batch

pytestmark = pytest.mark.gpu

TOP_K, KEEP, CONF_THR, OBJ_THR = 1000, 500, 0.01, 0.01


@pytest.fixture(scope='module')
def rd():
    import refinedet.pytorch_b200 as rd
    rd._ffi.lib()
    return rd


def _bench_seed(rank=0, buf=0, config=3):
    return 1234 + 1000 * config + rank + 100 * buf          # bench.py seed_for


def _run(rd, inputs, priors, size, C, nms_thr, instance=None, top_k=TOP_K, keep=KEEP, what='', repeat=1):
    arm_loc, arm_conf, odm_loc, odm_conf = inputs
    B = arm_loc.shape[0]
    det = rd.Detect_RefineDet(C, int(size), 0, top_k, CONF_THR, nms_thr, OBJ_THR, keep)
    scale = np.array([float(size)] * 4, np.float32)
    d_in = [t.cuda() for t in inputs]
    pri = priors.cuda()
    res = det.detect(*d_in, pri, scale=scale, instance=instance)
    # the boxes the stage used (same device function as a3 -> same bits); a3 itself against the oracle
    conf_a3 = odm_conf.clone().cuda()
    g_boxes, g_scores = det.forward(d_in[0], d_in[1], d_in[2], conf_a3, pri)
    o_boxes, o_scores = bo.detect_forward(arm_loc.numpy(), arm_conf.numpy(), odm_loc.numpy(), odm_conf.numpy().copy(),
                                          priors.numpy(), OBJ_THR)
    gb = g_boxes.cpu().numpy()
    np.testing.assert_allclose(gb, o_boxes, rtol=1e-5, atol=1e-6)          # north_star: 1e-5 relative
    assert np.array_equal(g_scores.cpu().numpy(), o_scores)
    counts, anchors, rows = detect_stage_eval_batch(gb, o_scores, scale, CONF_THR, top_k, nms_thr, keep)
    kept = assert_detections_equal(res, counts, anchors, rows, what)
    assert kept > 0
    for _ in range(repeat - 1):                                            # the same call again, again, ...
        res2 = det.detect(*d_in, pri, scale=scale, instance=instance)
        assert torch.equal(res2.counts, res.counts)
        m = torch.arange(res.dets.shape[2], device='cuda').view(1, 1, -1) < res.counts.unsqueeze(-1)
        assert torch.equal(res2.anchors[m], res.anchors[m]) and torch.equal(res2.dets[m], res.dets[m])
    return counts, int((arm_conf[..., 1] > OBJ_THR).sum()) // B


def test_headline_cfg3_bench_input_all_problems(rd):
    """bench.py's first input set (seed_for(0, 0), sparse generator): 32 x 81 problems against the oracle."""
    priors = rd.PriorBox(rd.REFINEDET_ANCHORS['512']).forward()
    inputs = gen.detect_inputs(_bench_seed(), 32, priors.shape[0], 81, 'sparse')
    counts, nodes = _run(rd, inputs, priors, 512, 81, 0.45, what='cfg3 sparse')
    assert 500 < nodes < 1024 and counts.max() <= 256            # the <256,128> regime the bench measures


@pytest.mark.parametrize('buf', [1, 5])
def test_headline_cfg3_other_bench_sets(rd, buf):
    priors = rd.PriorBox(rd.REFINEDET_ANCHORS['512']).forward()
    inputs = gen.detect_inputs(_bench_seed(0, buf), 32, priors.shape[0], 81, 'sparse')
    _run(rd, inputs, priors, 512, 81, 0.45, what='cfg3 sparse buf %d' % buf)


def test_headline_cfg3_dense_all_problems(rd):
    """The stress generator at full size: 14 k nodes per image (no graph), every class saturates top_k:
    radix select + per-problem bins in the queued large kernel, 2,560 problems."""
    priors = rd.PriorBox(rd.REFINEDET_ANCHORS['512']).forward()
    inputs = gen.detect_inputs(_bench_seed(0, 9), 32, priors.shape[0], 81, 'dense')
    _run(rd, inputs, priors, 512, 81, 0.45, what='cfg3 dense')


@pytest.mark.parametrize('n_obj', [6, 12, 24])
def test_headline_cfg3_clustered_all_problems(rd, n_obj):
    """Trained-detector-like inputs at B = 32, C = 81: dozens of mutually overlapping boxes per object ->
    images flagged wide-degree -> the <256,128> instance resolves through its triangular bit rows."""
    priors = rd.PriorBox(rd.REFINEDET_ANCHORS['512']).forward()
    inputs = gen.detect_inputs_clustered(900 + n_obj, 32, priors, 81, n_obj=n_obj)
    _run(rd, inputs, priors, 512, 81, 0.45, what='cfg3 clustered %d' % n_obj, repeat=3)


@pytest.mark.parametrize('kind', ['sparse', 'dense'])
def test_cfg2_refinedet320_voc_b32(rd, kind):
    priors = rd.PriorBox(rd.REFINEDET_ANCHORS['320']).forward()
    inputs = gen.detect_inputs(_bench_seed(0, 20, config=2), 32, priors.shape[0], 21, kind)
    _run(rd, inputs, priors, 320, 21, 0.45, what='cfg2 ' + kind)


@pytest.mark.parametrize('kind', ['sparse', 'dense'])
def test_cfg5_sarship_two_class_b32(rd, kind):
    """C = 2: B (C - 1) = 32 problems -> the <1024,256> instance (sparse: ~435 candidates per problem);
    dense: 12 k candidates per problem, radix select over the column."""
    priors = rd.PriorBox(rd.REFINEDET_ANCHORS['512']).forward()
    inputs = gen.detect_inputs(_bench_seed(0, 20, config=5), 32, priors.shape[0], 2, kind)
    _run(rd, inputs, priors, 512, 2, 0.49, what='cfg5 ' + kind)


def test_cfg5_clustered_b32(rd):
    priors = rd.PriorBox(rd.REFINEDET_ANCHORS['512']).forward()
    inputs = gen.detect_inputs_clustered(77, 32, priors, 2, n_obj=40)
    _run(rd, inputs, priors, 512, 2, 0.49, what='cfg5 clustered')


@pytest.mark.parametrize('instance', [256, 1024])
@pytest.mark.parametrize('kind,B,size,C,arm_shift', [
    ('sparse', 2, '512', 81, -8.0),       # ~700 nodes per image
    ('sparse', 4, '512', 2, -8.0),        # one class holds most nodes: 256 overflows -> queue, 1024 resolves
    ('sparse', 2, '512', 81, -7.0),       # ~1900 nodes: two-block graph
    ('sparse', 2, '320', 21, -5.0),       # ~2700 nodes of 6375
    ('clustered', 2, '512', 81, 0.0),     # wide-degree images: bit rows in 256, bin path for the others
    ('clustered', 3, '512', 2, 0.0),
    ('clustered', 2, '320', 21, 0.0),
    ('dense', 2, '320', 21, 0.0),         # no graph at all: every problem queued, whatever the instance
])
def test_forced_instances_vs_oracle(rd, instance, kind, B, size, C, arm_shift):
    """Every instantiation of the per-class kernel on every regime, whatever the heuristic would pick."""
    priors = rd.PriorBox(rd.REFINEDET_ANCHORS[size]).forward()
    if kind == 'clustered':
        inputs = gen.detect_inputs_clustered(4242 + B + C, B, priors, C)
    else:
        inputs = gen.detect_inputs(4242 + B + C, B, priors.shape[0], C, kind, arm_shift=arm_shift)
    _run(rd, inputs, priors, size, C, 0.45, instance=instance, what='%s forced %d' % (kind, instance))


@pytest.mark.parametrize('arm_shift', [-7.5, -7.0, -6.5, -6.0])
def test_density_sweep_b8(rd, arm_shift):
    """The node densities between the sparse and the dense generator (1.2 k - 4 k nodes per image, 257 - 1024
    candidates per problem at the upper end): multi-block graph, the mid-size instance, the queued graph resolve."""
    priors = rd.PriorBox(rd.REFINEDET_ANCHORS['512']).forward()
    inputs = gen.detect_inputs(31 + int(-10 * arm_shift), 8, priors.shape[0], 81, 'sparse', arm_shift=arm_shift)
    _run(rd, inputs, priors, 512, 81, 0.45, what='density %.1f' % arm_shift)


def test_graph_pair_list_drain_is_deterministic(rd):
    """graph_kernel drains its shared pair list between rounds of items once it is half full; the decision is
    CTA-uniform (__syncthreads_or).  Dense neighbourhoods (24 - 40 objects per image, > 1024 bin-surviving pairs per
    CTA) exercise the drain; 20 repeated runs must all equal the oracle-checked first one."""
    priors = rd.PriorBox(rd.REFINEDET_ANCHORS['512']).forward()
    for n_obj, C in ((24, 81), (40, 21)):
        inputs = gen.detect_inputs_clustered(5000 + n_obj, 8, priors, C, n_obj=n_obj, near_iou=0.3)
        _run(rd, inputs, priors, 512, C, 0.45, what='drain %d' % n_obj, repeat=20)


def test_negative_conf_thresh_admits_filtered_anchors(rd):
    """conf_thresh < 0: the reference zeroes the scores of ARM-filtered anchors (detection_refinedet.py:40-42)
    and then tests `score > conf_thresh` (eval_refinedet_coco.py:214), so they ARE candidates, with score 0."""
    priors = rd.PriorBox(rd.REFINEDET_ANCHORS['320']).forward()[::5].contiguous()
    P, C, B = priors.shape[0], 4, 2
    arm_loc, arm_conf, odm_loc, odm_conf = gen.detect_inputs(8, B, P, C, 'sparse', arm_shift=-3.0)
    det = rd.Detect_RefineDet(C, 320, 0, 2000, -0.5, 0.45, OBJ_THR, 2000)     # top_k > P: every anchor is a candidate
    scale = np.array([320.0] * 4, np.float32)
    d_in = [t.cuda() for t in (arm_loc, arm_conf, odm_loc, odm_conf)]
    res = det.detect(*d_in, priors.cuda(), scale=scale)
    g_boxes, g_scores = det.forward(d_in[0], d_in[1], d_in[2], odm_conf.clone().cuda(), priors.cuda())
    gs = g_scores.cpu().numpy()
    assert (gs.sum(-1) == 0).any() and (gs.sum(-1) > 0).any()        # both kinds of anchor present
    counts, anchors, rows = detect_stage_eval_batch(g_boxes.cpu().numpy(), gs, scale, -0.5, 2000, 0.45, 2000, workers=1)
    assert_detections_equal(res, counts, anchors, rows, 'conf_thresh < 0')
    assert (res.dets.cpu().numpy()[..., 4][np.arange(2000)[None, None] < counts[..., None]] == 0).any()
