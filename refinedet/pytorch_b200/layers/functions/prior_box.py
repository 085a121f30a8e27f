"""Anchor generator — the INPUT CONTRACT of the hot path (reference
``layers/functions/prior_box.py:28-56`` + the ``voc_refinedet`` / ``coco_refinedet`` dicts of
``data/config.py:63-119``).  Not accelerated: it runs once at model build.  Kept here so tests,
bench and smoke can produce the exact prior layout (level, row, col, anchor) without the
reference checkout."""
from math import sqrt

import numpy as np
import torch

# data/config.py:63-119 — both dicts share the anchor geometry
REFINEDET_ANCHORS = {
    '320': dict(feature_maps=[40, 20, 10, 5], min_dim=320, steps=[8, 16, 32, 64],
                min_sizes=[32, 64, 128, 256], max_sizes=[], aspect_ratios=[[2], [2], [2], [2]],
                variance=[0.1, 0.2], clip=True, name='RefineDet_320'),
    '512': dict(feature_maps=[64, 32, 16, 8], min_dim=512, steps=[8, 16, 32, 64],
                min_sizes=[32, 64, 128, 256], max_sizes=[], aspect_ratios=[[2], [2], [2], [2]],
                variance=[0.1, 0.2], clip=True, name='RefineDet_512'),
}


class PriorBox(object):
    """``PriorBox(cfg).forward()`` -> ``[P,4]`` (cx,cy,w,h) float32, like the reference."""

    def __init__(self, cfg):
        self.cfg = cfg
        for v in (cfg.get('variance') or [0.1]):
            if v <= 0:
                raise ValueError('Variances must be greater than 0')

    def forward(self):
        cfg = self.cfg
        levels = []
        for k, f in enumerate(cfg['feature_maps']):
            f_k = cfg['min_dim'] / cfg['steps'][k]
            s_k = cfg['min_sizes'][k] / cfg['min_dim']
            shapes = [(s_k, s_k)]
            if cfg['max_sizes']:
                s_p = sqrt(s_k * (cfg['max_sizes'][k] / cfg['min_dim']))
                shapes.append((s_p, s_p))
            for ar in cfg['aspect_ratios'][k]:
                shapes.append((s_k * sqrt(ar), s_k / sqrt(ar)))
                shapes.append((s_k / sqrt(ar), s_k * sqrt(ar)))
            centre = (np.arange(f, dtype=np.float64) + 0.5) / f_k
            cy, cx = np.meshgrid(centre, centre, indexing='ij')          # row i -> cy, col j -> cx
            wh = np.asarray(shapes, dtype=np.float64)                     # [A,2]
            lvl = np.empty((f, f, len(shapes), 4), dtype=np.float64)
            lvl[..., 0] = cx[..., None]
            lvl[..., 1] = cy[..., None]
            lvl[..., 2] = wh[:, 0]
            lvl[..., 3] = wh[:, 1]
            levels.append(lvl.reshape(-1, 4))
        out = torch.from_numpy(np.concatenate(levels, 0).astype(np.float32))
        if cfg['clip']:
            out.clamp_(max=1, min=0)
        return out
