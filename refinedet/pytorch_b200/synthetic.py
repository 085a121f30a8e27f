"""Seeded synthetic inputs for the detect / match hot path (SURVEY.md §8d generators).

Generated on the CPU with a seeded ``torch.Generator`` so the oracle and the kernels see the
same bits.  Used by tests, ``bench.py`` and ``__graft_entry__.smoke()``."""
import numpy as np
import torch


def detect_logits(seed, B, P, C, kind='sparse', arm_shift=-8.0):
    """Returns arm_loc[B,P,4], arm_logits[B,P,2], odm_loc[B,P,4], odm_logits[B,P,C]: the head outputs
    BEFORE the softmax of models/refinedet.py:143-147 (the ``RD_INPUT_LOGITS`` form of the stage).

    ``dense``  : loc = 0.5 randn, arm/odm logits = 3 randn                 (stress: ~86 % pass ARM)
    ``sparse`` : loc = randn, ARM logit gap 2 randn + arm_shift, ODM logits 1.5 randn with +4 on
                 class 0                                                   (realistic: ~4 % pass ARM)
    """
    g = torch.Generator().manual_seed(int(seed))
    if kind == 'sparse':
        d = 2.0 * torch.randn(B, P, generator=g) + arm_shift
        arm_logits = torch.stack([torch.zeros(B, P), d], -1)
        odm_logits = 1.5 * torch.randn(B, P, C, generator=g)
        odm_logits[..., 0] += 4.0
        loc_s = 1.0
    elif kind == 'dense':
        arm_logits = 3 * torch.randn(B, P, 2, generator=g)
        odm_logits = 3 * torch.randn(B, P, C, generator=g)
        loc_s = 0.5
    else:
        raise ValueError(kind)
    arm_loc = loc_s * torch.randn(B, P, 4, generator=g)
    odm_loc = loc_s * torch.randn(B, P, 4, generator=g)
    return arm_loc.contiguous(), arm_logits.contiguous(), odm_loc.contiguous(), odm_logits.contiguous()


def detect_inputs(seed, B, P, C, kind='sparse', arm_shift=-8.0):
    """Returns arm_loc[B,P,4], arm_conf[B,P,2], odm_loc[B,P,4], odm_conf[B,P,C] (softmaxed):
    :func:`detect_logits` followed by the model's softmax."""
    arm_loc, arm_logits, odm_loc, odm_logits = detect_logits(seed, B, P, C, kind, arm_shift)
    return (arm_loc, torch.softmax(arm_logits, -1).contiguous(), odm_loc,
            torch.softmax(odm_logits, -1).contiguous())


def targets(seed, B, G, num_classes, wh_lo=0.02, wh_hi=0.17):
    """list of B tensors [G,5]: xy ~ U(0,0.8), wh ~ U(wh_lo,wh_hi), label randint(1,num_classes)."""
    g = torch.Generator().manual_seed(int(seed))
    out = []
    for _ in range(B):
        xy = torch.rand(G, 2, generator=g) * 0.8
        wh = wh_lo + torch.rand(G, 2, generator=g) * (wh_hi - wh_lo)
        x2y2 = torch.clamp(xy + wh, max=1.0)
        lab = torch.randint(1, max(num_classes, 2), (G, 1), generator=g).float()
        out.append(torch.cat([xy, x2y2, lab], 1))
    return out


def train_predictions(seed, B, P, C):
    """cfg 4: logits (randn) for both heads, loc = 0.1 randn."""
    g = torch.Generator().manual_seed(int(seed))
    arm_loc = 0.1 * torch.randn(B, P, 4, generator=g)
    odm_loc = 0.1 * torch.randn(B, P, 4, generator=g)
    arm_conf = torch.randn(B, P, 2, generator=g)
    odm_conf = torch.randn(B, P, C, generator=g)
    return arm_loc, arm_conf, odm_loc, odm_conf


def assert_tie_free(scores, conf_thresh):
    """per (image,class) candidate scores pairwise distinct (SURVEY.md §8d tie-freedom)."""
    s = scores.numpy() if isinstance(scores, torch.Tensor) else scores
    B, P, C = s.shape
    for b in range(B):
        for c in range(1, C):
            v = s[b, :, c]
            v = v[v > np.float32(conf_thresh)]
            if np.unique(v).size != v.size:
                return False
    return True


def tie_free_detect_inputs(seed, B, P, C, kind, conf_thresh, obj_thresh, arm_shift=-8.0, max_tries=20):
    """``detect_inputs`` with the SURVEY.md §8d guarantee: per (image, class) the candidate scores
    (ARM-passing anchors with score > conf_thresh) are pairwise distinct; ``seed + 1`` otherwise."""
    for k in range(max_tries):
        a = detect_inputs(seed + k, B, P, C, kind, arm_shift=arm_shift)
        scores = a[3].clone()
        scores[a[1][..., 1] <= obj_thresh] = 0
        if assert_tie_free(scores, conf_thresh):
            return a
    raise RuntimeError('no tie-free input found in %d seeds' % max_tries)
