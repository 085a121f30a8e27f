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


def detect_logits_clustered(seed, B, priors, C, n_obj=12, background_shift=-9.0, jitter=1.0, near_iou=0.35):
    """Head outputs shaped like a TRAINED detector's: ``n_obj`` objects per image, every anchor whose prior
    overlaps an object (IoU > 0.35) passes the ARM gate, regresses to that object with jitter and scores its
    class -- dozens of mutually overlapping boxes per object, the case greedy NMS exists for -- over the
    ``sparse`` background (ARM logit gap 2 randn + ``background_shift``).  ``priors``: [P,4] (cx,cy,w,h).
    ``jitter`` scales the regression noise: small values make the boxes of an object nearly coincide, so that
    a node has as many suppressors as its object has anchors (beyond the 64 the suppression graph keeps)."""
    g = torch.Generator().manual_seed(int(seed))
    P = priors.shape[0]
    pr = priors.float().cpu()
    pxy1, pxy2 = pr[:, :2] - pr[:, 2:] / 2, pr[:, :2] + pr[:, 2:] / 2
    wh = 0.08 + 0.32 * torch.rand(B, n_obj, 2, generator=g)
    cxy = 0.1 + 0.8 * torch.rand(B, n_obj, 2, generator=g)
    cls = torch.randint(1, max(C, 2), (B, n_obj), generator=g)
    gx1, gx2 = cxy - wh / 2, cxy + wh / 2
    lt = torch.maximum(pxy1[None, :, None, :], gx1[:, None, :, :])
    rb = torch.minimum(pxy2[None, :, None, :], gx2[:, None, :, :])
    inter = (rb - lt).clamp(min=0).prod(-1)                                     # [B,P,n_obj]
    iou = inter / (pr[:, 2:].prod(-1)[None, :, None] + wh.prod(-1)[:, None, :] - inter)
    best, which = iou.max(-1)                                                   # [B,P]
    near = best > near_iou
    bi = torch.arange(B)[:, None]
    ocx, owh, ocl = cxy[bi, which], wh[bi, which], cls[bi, which]               # the anchor's object
    enc = torch.cat([(ocx - pr[None, :, :2]) / (0.1 * pr[None, :, 2:]),
                     torch.log(owh / pr[None, :, 2:]) / 0.2], -1)               # encode(object, prior)
    d = torch.where(near, 3.0 + torch.randn(B, P, generator=g), 2.0 * torch.randn(B, P, generator=g) + background_shift)
    arm_logits = torch.stack([torch.zeros(B, P), d], -1)
    arm_loc = torch.where(near[..., None], (1.0 - 0.2 * min(jitter, 1.0)) * enc + 0.3 * jitter * torch.randn(B, P, 4, generator=g),
                          torch.randn(B, P, 4, generator=g))
    odm_loc = torch.where(near[..., None], 0.4 * jitter * torch.randn(B, P, 4, generator=g), torch.randn(B, P, 4, generator=g))
    odm_logits = 1.5 * torch.randn(B, P, C, generator=g)
    odm_logits[..., 0] += 4.0
    boost = torch.zeros(B, P, C)
    boost.scatter_(2, ocl[..., None], 7.0)
    odm_logits = odm_logits + boost * near[..., None]
    return arm_loc.contiguous(), arm_logits.contiguous(), odm_loc.contiguous(), odm_logits.contiguous()


def detect_inputs_clustered(seed, B, priors, C, n_obj=12, jitter=1.0, near_iou=0.35):
    arm_loc, arm_logits, odm_loc, odm_logits = detect_logits_clustered(seed, B, priors, C, n_obj, jitter=jitter,
                                                                       near_iou=near_iou)
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
