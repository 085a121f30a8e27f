"""CPU tier: the reference arm of bench.py (baseline/reference_arm.py driving the UNMODIFIED reference staged under
baseline/_ref) computes the same detections and the same losses as the oracle — i.e. the two arms of the benchmark
measure the same work — and bench.py describes the workload identically in both arms."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from baseline import reference_arm as ra  # noqa: E402
from oracle import box_oracle as bo  # noqa: E402
from refinedet.pytorch_b200 import synthetic  # noqa: E402

needs_ref = pytest.mark.skipif(not ra.available(), reason='baseline/_ref not staged (run __graft_entry__.build() where '
                                                         '/root/reference exists)')


@needs_ref
@pytest.mark.filterwarnings('ignore')
@pytest.mark.parametrize('kind,size,C', [('sparse', '320', 21), ('dense', '320', 5)])
def test_reference_arm_detect_equals_oracle(kind, size, C):
    R = ra.load_reference()
    priors = R.PriorBox(R.voc[size]).forward()
    P = priors.shape[0]
    # tie-free candidates: the reference's argsort()[::-1] orders exactly tied scores differently from the documented rule
    a = synthetic.tie_free_detect_inputs(11, 1, P, C, kind, 0.01, 0.01, arm_shift=-3.0)
    det = R.Detect(C, int(size), 0, 200, 0.01, 0.45, 0.01, 100)
    scale = torch.tensor([float(size)] * 4)
    with torch.no_grad():
        out = ra.detect_one_image(R, det, a[0], a[1], a[2], a[3].clone(), priors, scale, 0.01, 200, 0.45, 100)
    boxes, scores = bo.detect_forward(a[0].numpy(), a[1].numpy(), a[2].numpy(), a[3].numpy().copy(), priors.numpy(), 0.01)
    exp, _ = bo.detect_stage_eval(boxes[0], scores[0], scale.numpy(), 0.01, 200, 0.45, 100)
    assert len(out) == len(exp) == C
    kept = 0
    for j in range(C):
        assert out[j].shape == exp[j].shape, j
        assert np.array_equal(out[j][:, 4], exp[j][:, 4])                       # scores: copies
        np.testing.assert_allclose(out[j][:, :4], exp[j][:, :4], rtol=5e-5, atol=1e-4)   # numpy vs torch exp through two decodes
        kept += out[j].shape[0]
    assert kept > 0


@needs_ref
@pytest.mark.filterwarnings('ignore')
def test_reference_arm_train_step_equals_oracle():
    B, C, G = 2, 5, 6
    R = ra.load_reference()
    P = 16320
    sec, vals = ra.run_train_step(3, 4, B, P, C, G, steps=1, threads=2)
    assert sec > 0 and len(vals) == 4 and all(np.isfinite(vals))
    tp = [t.numpy() for t in synthetic.train_predictions(3, B, P, C)]
    tg = [t.numpy() for t in synthetic.targets(4, B, G, C)]
    priors = bo.prior_box(bo.REFINEDET_CFG['512'])
    arm = bo.multibox_loss(tp + [priors], tg, 2, use_ARM=False)
    odm = bo.multibox_loss(tp + [priors], tg, C, use_ARM=True)
    np.testing.assert_allclose(vals, [arm['loss_l'], arm['loss_c'], odm['loss_l'], odm['loss_c']], rtol=2e-5)


def test_bench_config_is_the_same_in_both_arms():
    import importlib.util
    spec = importlib.util.spec_from_file_location('_bench', os.path.join(ROOT, 'bench.py'))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    for workload in ('sparse', 'dense'):
        for scaling in ('weak', 'strong'):
            c = bench.make_config(workload, scaling)
            assert c == bench.make_config(workload, scaling)
            assert set(c) == {'workload', 'batch_per_gpu', 'anchors', 'classes', 'generator', 'sharding'}
            assert 'P=16320' in c['workload'] and 'C=81' in c['workload']
    assert bench.seed_for(0, 0) == 4234
    assert bench.bind_to_gpu_numa_node(0) is None or bench.bind_to_gpu_numa_node(0) > 0
