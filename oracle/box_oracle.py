"""numpy restatement of the reference's detect / anchor-matching arithmetic.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  Every function cites the
reference ``file:line`` (relative to the reference checkout) it follows.  All
arithmetic is float32 in the reference's operation order; comparisons against
Python-float thresholds are done in float32 (probe: both ``torch.le`` and numpy
compare an f32 array against a Python scalar after rounding the scalar to f32).

Tie rule.  The reference's sorts (``torch.sort``, ``numpy.argsort``) are unstable,
so its result on exactly tied scores is implementation-defined.  This oracle —
and the CUDA path — define "score descending, lower index first".  Tests use
tie-free inputs, where the rule is unobservable.
"""
import numpy as np

F32 = np.float32


def _f(x):
    return np.asarray(x, dtype=F32)


# ----------------------------------------------------------------------------
# layers/box_utils.py
# ----------------------------------------------------------------------------
def point_form(boxes):
    """layers/box_utils.py:5-14  (cx,cy,w,h) -> (x1,y1,x2,y2)."""
    boxes = _f(boxes)
    half = boxes[:, 2:] / F32(2)
    return np.concatenate((boxes[:, :2] - half, boxes[:, :2] + half), 1)


def center_size(boxes):
    """layers/box_utils.py:17-26  (x1,y1,x2,y2) -> (cx,cy,w,h)."""
    boxes = _f(boxes)
    return np.concatenate(((boxes[:, 2:] + boxes[:, :2]) / F32(2),
                           boxes[:, 2:] - boxes[:, :2]), 1)


def intersect(box_a, box_b):
    """layers/box_utils.py:29-47  [A,4] x [B,4] -> [A,B] intersection area."""
    box_a, box_b = _f(box_a), _f(box_b)
    max_xy = np.minimum(box_a[:, None, 2:], box_b[None, :, 2:])
    min_xy = np.maximum(box_a[:, None, :2], box_b[None, :, :2])
    inter = np.maximum(max_xy - min_xy, F32(0))
    return inter[:, :, 0] * inter[:, :, 1]


def jaccard(box_a, box_b):
    """layers/box_utils.py:50-68  IoU matrix; union = (area_a + area_b) - inter."""
    box_a, box_b = _f(box_a), _f(box_b)
    inter = intersect(box_a, box_b)
    area_a = ((box_a[:, 2] - box_a[:, 0]) * (box_a[:, 3] - box_a[:, 1]))[:, None]
    area_b = ((box_b[:, 2] - box_b[:, 0]) * (box_b[:, 3] - box_b[:, 1]))[None, :]
    union = area_a + area_b - inter
    with np.errstate(divide='ignore', invalid='ignore'):
        return inter / union


def encode(matched, priors, variances):
    """layers/box_utils.py:162-183."""
    matched, priors = _f(matched), _f(priors)
    v0, v1 = F32(variances[0]), F32(variances[1])
    g_cxcy = (matched[:, :2] + matched[:, 2:]) / F32(2) - priors[:, :2]
    g_cxcy = g_cxcy / (v0 * priors[:, 2:])
    g_wh = (matched[:, 2:] - matched[:, :2]) / priors[:, 2:]
    with np.errstate(divide='ignore', invalid='ignore'):
        g_wh = np.log(g_wh + F32(1e-5)) / v1
    return np.concatenate([g_cxcy, g_wh], 1)


def decode(loc, priors, variances):
    """layers/box_utils.py:187-205.  Note x2 = w + x1 (not cx + w/2)."""
    loc, priors = _f(loc), _f(priors)
    v0, v1 = F32(variances[0]), F32(variances[1])
    cxcy = priors[:, :2] + (loc[:, :2] * v0) * priors[:, 2:]
    wh = priors[:, 2:] * np.exp(loc[:, 2:] * v1)
    x1y1 = cxcy - wh / F32(2)
    x2y2 = wh + x1y1
    return np.concatenate((x1y1, x2y2), 1)


def log_sum_exp(x):
    """layers/box_utils.py:208-216.  Max is over the WHOLE tensor (:215)."""
    x = _f(x)
    x_max = x.max()
    return np.log(np.sum(np.exp(x - x_max), 1, keepdims=True, dtype=F32)) + x_max


def _order_desc(scores):
    """Score descending, lower index first on ties (the documented tie rule)."""
    scores = _f(scores)
    return np.argsort(-scores, kind='stable')


def nms(boxes, scores, overlap=0.5, top_k=200):
    """layers/box_utils.py:222-286  greedy NMS on normalised boxes (no +1).

    Returns ``(keep[int64, n] zero padded, count)``; for empty input returns the
    bare ``keep`` array, mirroring :235-236.
    IoU = inter / ((area_j - inter) + area_i), keep candidates with IoU <= overlap.
    """
    boxes, scores = _f(boxes).reshape(-1, 4), _f(scores).reshape(-1)
    n = scores.shape[0]
    keep = np.zeros(n, dtype=np.int64)
    if boxes.size == 0:
        return keep
    x1, y1, x2, y2 = boxes[:, 0], boxes[:, 1], boxes[:, 2], boxes[:, 3]
    area = (x2 - x1) * (y2 - y1)
    idx = _order_desc(scores)[:top_k]          # :242-244 (top_k largest), descending here
    thr = F32(overlap)
    count = 0
    while idx.size > 0:
        i = idx[0]
        keep[count] = i
        count += 1
        if idx.size == 1:
            break
        idx = idx[1:]
        xx1 = np.maximum(x1[idx], x1[i])
        yy1 = np.maximum(y1[idx], y1[i])
        xx2 = np.minimum(x2[idx], x2[i])
        yy2 = np.minimum(y2[idx], y2[i])
        w = np.maximum(xx2 - xx1, F32(0))
        h = np.maximum(yy2 - yy1, F32(0))
        inter = w * h
        union = (area[idx] - inter) + area[i]
        with np.errstate(divide='ignore', invalid='ignore'):
            iou = inter / union
        idx = idx[iou <= thr]
    return keep, count


# ----------------------------------------------------------------------------
# utils/nms/*  (pixel coordinates, +1 convention)
# ----------------------------------------------------------------------------
def nms_pixel(dets, thresh, suppress_on_equal=False):
    """utils/nms/py_cpu_nms.py:10-38 (== nms_kernel.cu:24-32,71 semantics).

    ``dets[n,5]`` = x1,y1,x2,y2,score in pixels.  Suppress when IoU > thresh;
    with ``suppress_on_equal`` when IoU >= thresh (utils/nms/cpu_nms.pyx:65).
    Returns the list of kept row indices, score-descending.
    """
    dets = _f(dets)
    if dets.shape[0] == 0:
        return []                               # utils/nms_wrapper.py:26-27
    x1, y1, x2, y2, scores = (dets[:, k] for k in range(5))
    areas = (x2 - x1 + F32(1)) * (y2 - y1 + F32(1))
    order = _order_desc(scores)
    thr = F32(thresh)
    keep = []
    while order.size > 0:
        i = order[0]
        keep.append(int(i))
        rest = order[1:]
        xx1 = np.maximum(x1[i], x1[rest])
        yy1 = np.maximum(y1[i], y1[rest])
        xx2 = np.minimum(x2[i], x2[rest])
        yy2 = np.minimum(y2[i], y2[rest])
        w = np.maximum(F32(0), xx2 - xx1 + F32(1))
        h = np.maximum(F32(0), yy2 - yy1 + F32(1))
        inter = w * h
        ovr = inter / (areas[i] + areas[rest] - inter)
        if suppress_on_equal:
            order = rest[ovr < thr]
        else:
            order = rest[ovr <= thr]
    return keep


# ----------------------------------------------------------------------------
# layers/functions/detection_refinedet.py
# ----------------------------------------------------------------------------
def detect_forward(arm_loc, arm_conf, odm_loc, odm_conf, priors,
                   objectness_thre=0.01, variance=(0.1, 0.2)):
    """Detect_RefineDet.forward, layers/functions/detection_refinedet.py:27-65.

    Mutates ``odm_conf`` in place (:40-42): rows whose ARM objectness
    ``arm_conf[...,1] <= objectness_thre`` become all-zero (background column
    included).  Returns ``(boxes[B,P,4], scores[B,P,C])`` — fresh arrays.
    """
    arm_loc, arm_conf, odm_loc, priors = _f(arm_loc), _f(arm_conf), _f(odm_loc), _f(priors)
    assert odm_conf.dtype == F32
    no_obj = arm_conf[:, :, 1] <= F32(objectness_thre)
    odm_conf[no_obj] = F32(0)
    B, P = odm_loc.shape[:2]
    boxes = np.zeros((B, P, 4), dtype=F32)
    for i in range(B):
        default = center_size(decode(arm_loc[i], priors, variance))   # :57-58
        boxes[i] = decode(odm_loc[i], default, variance)              # :59
    return boxes, odm_conf.copy()


def detect_stage_eval(boxes, scores, scale, conf_thresh=0.01, top_k=1000,
                      nms_thresh=0.45, max_per_image=500, suppress_on_equal=False):
    """The detect stage as evaluated: eval_refinedet_coco.py:205-232, one image.

    ``boxes[P,4]`` normalised, ``scores[P,C]``, ``scale[4]`` pixels.  Returns a list
    of length C; entry 0 is an empty array (background is never evaluated, :213),
    entry j is ``[n_j,5]`` (x1,y1,x2,y2,score) in pixels, score-descending, and
    ``kept_anchor`` lists the anchor index of every row.
    """
    boxes = _f(boxes) * _f(scale)[None, :]                    # :209
    scores = _f(scores)
    C = scores.shape[1]
    out = [np.empty((0, 5), dtype=F32)]
    anchors = [np.empty((0,), dtype=np.int64)]
    for j in range(1, C):
        inds = np.where(scores[:, j] > F32(conf_thresh))[0]   # :214
        if len(inds) == 0:
            out.append(np.empty((0, 5), dtype=F32))
            anchors.append(np.empty((0,), dtype=np.int64))
            continue
        c_scores = scores[inds, j]
        order = _order_desc(c_scores)[:top_k]                 # :222
        c_dets = np.hstack((boxes[inds][order], c_scores[order][:, None])).astype(F32)
        keep = nms_pixel(c_dets, nms_thresh, suppress_on_equal)   # :229
        keep = keep[:max_per_image]                           # :231
        out.append(c_dets[keep, :])
        anchors.append(inds[order][keep])
    return out, anchors


def select_topk(scores, conf_thresh, top_k, first_class=1):
    """The candidate list the reference hands to NMS, per (image, class): ``where(scores[:, j] > thresh)``
    (eval_refinedet_coco.py:214; detection_refinedet.py:98) then ``argsort()[::-1][:top_k]`` (eval :222;
    the last ``top_k`` of the ascending sort in box_utils.py:242-244), lower anchor first on ties.

    ``scores[B,P,C]`` -> ``lists[b][c]`` = int64 anchor indices, score-descending (empty for c < first_class)."""
    scores = _f(scores)
    B, P, C = scores.shape
    out = []
    for b in range(B):
        row = []
        for c in range(C):
            if c < first_class:
                row.append(np.empty((0,), dtype=np.int64))
                continue
            inds = np.where(scores[b, :, c] > F32(conf_thresh))[0]
            order = _order_desc(scores[b, inds, c])[:top_k]
            row.append(inds[order].astype(np.int64))
        out.append(row)
    return out


def forward_python_nms(arm_loc, arm_conf, odm_loc, odm_conf, priors, num_classes,
                       top_k, conf_thresh, nms_thresh, objectness_thre=0.01,
                       variance=(0.1, 0.2)):
    """Detect_RefineDet.forward_python_nms, detection_refinedet.py:67-113.

    Returns ``output[B,C,top_k,5]`` rows (score,x1,y1,x2,y2), zero padded, class 0
    all zero.  The cross-class keep_top_k step (:109-112) is a no-op in the
    reference (it fills a temporary) and is therefore absent here.
    """
    boxes, scores = detect_forward(arm_loc, arm_conf, odm_loc, odm_conf, priors,
                                   objectness_thre, variance)
    B, P, C = scores.shape
    assert C == num_classes
    output = np.zeros((B, C, top_k, 5), dtype=F32)
    anchors = -np.ones((B, C, top_k), dtype=np.int64)
    for i in range(B):
        for cl in range(1, C):
            c_mask = scores[i, :, cl] > F32(conf_thresh)      # :98
            sc = scores[i, c_mask, cl]
            if sc.shape[0] == 0:
                continue
            bx = boxes[i][c_mask]
            ids, count = nms(bx, sc, nms_thresh, top_k)       # :105
            ids = ids[:count]
            output[i, cl, :count, 0] = sc[ids]
            output[i, cl, :count, 1:] = bx[ids]
            anchors[i, cl, :count] = np.where(c_mask)[0][ids]
    return output, anchors


def coco_results(all_boxes, image_ids, class_to_cat_id):
    """data/sarship_coco.py:293-336 (``_coco_results_one_category`` + ``_write_coco_results_file``):
    ``all_boxes[c][i]`` = ``[n,5]`` rows (x1,y1,x2,y2,score) -> list of result dicts, classes ascending
    (class 0 = background skipped), images ascending; bbox = [x, y, x2-x+1, y2-y+1] in float64."""
    results = []
    for c in range(1, len(all_boxes)):
        for i, index in enumerate(image_ids):
            dets = np.asarray(all_boxes[c][i]).astype(np.float64)      # :296 astype(np.float)
            if dets.shape[0] == 0:                                     # :297-298
                continue
            scores = dets[:, -1]
            xs, ys = dets[:, 0], dets[:, 1]
            ws = dets[:, 2] - xs + 1                                   # :302
            hs = dets[:, 3] - ys + 1
            results.extend({'image_id': index, 'category_id': class_to_cat_id[c],
                            'bbox': [xs[k], ys[k], ws[k], hs[k]], 'score': scores[k]}
                           for k in range(dets.shape[0]))
    return results


# ----------------------------------------------------------------------------
# layers/box_utils.py: match / refine_match
# ----------------------------------------------------------------------------
def _first_argmax(a, axis):
    # numpy argmax returns the first maximal index, like torch.max on CPU
    return np.argmax(a, axis=axis)


def refine_match(threshold, truths, priors, variances, labels, arm_loc=None,
                 label_offset=0):
    """layers/box_utils.py:113-160 (and ``match`` :70-111 with ``label_offset=1``).

    Returns ``(loc[P,4] f32, conf[P] int64, best_truth_idx[P], best_truth_overlap[P])``.
    ``labels`` may be float (ODM: 1-based class) or bool (ARM 2-class,
    refinedet_multibox_loss.py:78-79).
    """
    truths, priors = _f(truths), _f(priors)
    labels = np.asarray(labels)
    if arm_loc is None:
        anchors_pt = point_form(priors)                       # :133
    else:
        anchors_pt = decode(arm_loc, priors, variances)       # :135
    overlaps = jaccard(truths, anchors_pt)                    # [G,P]
    best_prior_idx = _first_argmax(overlaps, 1)               # :139
    best_truth_idx = _first_argmax(overlaps, 0)               # :141
    best_truth_overlap = overlaps[best_truth_idx, np.arange(overlaps.shape[1])].copy()
    best_truth_overlap[best_prior_idx] = F32(2)               # :146
    for j in range(best_prior_idx.shape[0]):                  # :149-150 (last j wins)
        best_truth_idx[best_prior_idx[j]] = j
    matches = truths[best_truth_idx]                          # :151
    conf = labels[best_truth_idx]
    if label_offset:
        conf = conf + label_offset                            # match(): labels + 1, :107
    if arm_loc is None:
        loc = encode(matches, priors, variances)              # :154
    else:
        loc = encode(matches, center_size(anchors_pt), variances)   # :157
    conf = conf.copy()
    conf[best_truth_overlap < F32(threshold)] = 0             # :158
    return loc, conf.astype(np.int64), best_truth_idx, best_truth_overlap


# ----------------------------------------------------------------------------
# layers/modules/refinedet_multibox_loss.py
# ----------------------------------------------------------------------------
def hnm_select(loss_c, pos, negpos_ratio=3):
    """refinedet_multibox_loss.py:117-123.

    ``loss_c[B,P]`` (this function zeroes the positives itself, :117), ``pos[B,P]``
    bool.  ``neg = rank < min(ratio*num_pos, P-1)`` with rank from a descending
    sort — ties ranked lower index first.
    """
    loss_c = _f(loss_c).copy()
    pos = np.asarray(pos, dtype=bool)
    loss_c[pos] = F32(0)
    B, P = loss_c.shape
    num_pos = pos.sum(1)
    num_neg = np.minimum(negpos_ratio * num_pos, P - 1)
    neg = np.zeros((B, P), dtype=bool)
    for b in range(B):
        order = _order_desc(loss_c[b])
        neg[b, order[:num_neg[b]]] = True
    return neg, loss_c


def softmax(x, axis=-1):
    x = _f(x)
    m = x.max(axis=axis, keepdims=True)
    e = np.exp(x - m)
    return e / e.sum(axis=axis, keepdims=True, dtype=F32)


def multibox_loss(predictions, targets, num_classes, overlap_thresh=0.5,
                  negpos_ratio=3, theta=0.01, use_ARM=False, variance=(0.1, 0.2)):
    """RefineDetMultiBoxLoss.forward, refinedet_multibox_loss.py:50-139 (values only).

    Returns a dict with ``loss_l``, ``loss_c`` (float), and the intermediate targets
    and masks (``loc_t, conf_t, pos, neg``).  Sums are accumulated in float64 and
    compared with tolerance by the tests (torch's reduction order is not specified).
    """
    arm_loc, arm_conf, odm_loc, odm_conf, priors = [_f(t) for t in predictions]
    loc_data, conf_data = (odm_loc, odm_conf) if use_ARM else (arm_loc, arm_conf)
    B, P = loc_data.shape[:2]
    priors = priors[:P]                                        # :68
    loc_t = np.zeros((B, P, 4), dtype=F32)
    conf_t = np.zeros((B, P), dtype=np.int64)
    for idx in range(B):
        t = _f(targets[idx])
        truths, labels = t[:, :-1], t[:, -1]
        if num_classes == 2 and not use_ARM:
            labels = labels >= 0                               # :78-79
        l, c, _, _ = refine_match(overlap_thresh, truths, priors, variance, labels,
                                  arm_loc[idx] if use_ARM else None)
        loc_t[idx], conf_t[idx] = l, c
    pos = conf_t > 0
    if use_ARM:                                                # :96-101
        p_obj = softmax(arm_conf, 2)[:, :, 1]
        pos = pos & ~(p_obj <= F32(theta))
    d = (loc_data[pos] - loc_t[pos]).astype(np.float64)        # :107-110 SmoothL1(sum), beta=1
    ad = np.abs(d)
    loss_l = np.where(ad < 1.0, 0.5 * d * d, ad - 0.5).sum()
    batch_conf = conf_data.reshape(-1, num_classes)
    lse = log_sum_exp(batch_conf)[:, 0]
    loss_c_all = (lse - batch_conf[np.arange(B * P), conf_t.reshape(-1)]).reshape(B, P)   # :114
    neg, loss_c_mined = hnm_select(loss_c_all, pos, negpos_ratio)
    sel = pos | neg
    x = conf_data[sel].astype(np.float64)                      # :126-130 CE(sum)
    t_sel = conf_t[sel]
    m = x.max(1, keepdims=True)
    lse2 = np.log(np.exp(x - m).sum(1)) + m[:, 0]
    loss_c = (lse2 - x[np.arange(x.shape[0]), t_sel]).sum()
    N = float(pos.sum())
    if N < 1:                                                  # :135-136
        loss_l, loss_c = 0.0, 0.0
    else:
        loss_l, loss_c = loss_l / N, loss_c / N
    return dict(loss_l=loss_l, loss_c=loss_c, loc_t=loc_t, conf_t=conf_t, pos=pos, neg=neg,
                loss_c_rows=loss_c_mined, N=N)


def multibox_loss_grads(loc_data, conf_data, loc_t, conf_t, pos, neg, N):
    """Gradients autograd derives from refinedet_multibox_loss.py:105-138 for
    ``loss_l + loss_c`` (both divided by N): SmoothL1'(d) = clamp(d, -1, 1) on the positives,
    softmax(x) - onehot(t) on ``pos | neg``, zero elsewhere.  float64 inside, returned as float32."""
    loc_data, conf_data, loc_t = _f(loc_data), _f(conf_data), _f(loc_t)
    g_loc = np.zeros(loc_data.shape, np.float64)
    g_conf = np.zeros(conf_data.shape, np.float64)
    if N >= 1:
        d = (loc_data[pos] - loc_t[pos]).astype(np.float64)
        g_loc[pos] = np.clip(d, -1.0, 1.0) / N
        sel = pos | neg
        x = conf_data[sel].astype(np.float64)
        e = np.exp(x - x.max(1, keepdims=True))
        sm = e / e.sum(1, keepdims=True)
        sm[np.arange(sm.shape[0]), conf_t[sel]] -= 1.0
        g_conf[sel] = sm / N
    return g_loc.astype(F32), g_conf.astype(F32)


# ----------------------------------------------------------------------------
# layers/functions/prior_box.py + data/config.py (input contract a0)
# ----------------------------------------------------------------------------
REFINEDET_CFG = {
    # data/config.py:63-119 — voc_refinedet / coco_refinedet share the anchor geometry
    '320': dict(feature_maps=[40, 20, 10, 5], min_dim=320, steps=[8, 16, 32, 64],
                min_sizes=[32, 64, 128, 256], max_sizes=[], aspect_ratios=[[2], [2], [2], [2]],
                variance=[0.1, 0.2], clip=True),
    '512': dict(feature_maps=[64, 32, 16, 8], min_dim=512, steps=[8, 16, 32, 64],
                min_sizes=[32, 64, 128, 256], max_sizes=[], aspect_ratios=[[2], [2], [2], [2]],
                variance=[0.1, 0.2], clip=True),
}


def prior_box(cfg):
    """PriorBox.forward, layers/functions/prior_box.py:28-56 (Python-float math, cast to f32)."""
    from math import sqrt
    mean = []
    for k, f in enumerate(cfg['feature_maps']):
        for i in range(f):
            for j in range(f):
                f_k = cfg['min_dim'] / cfg['steps'][k]
                cx = (j + 0.5) / f_k
                cy = (i + 0.5) / f_k
                s_k = cfg['min_sizes'][k] / cfg['min_dim']
                mean += [cx, cy, s_k, s_k]
                if cfg['max_sizes']:
                    s_p = sqrt(s_k * (cfg['max_sizes'][k] / cfg['min_dim']))
                    mean += [cx, cy, s_p, s_p]
                for ar in cfg['aspect_ratios'][k]:
                    mean += [cx, cy, s_k * sqrt(ar), s_k / sqrt(ar)]
                    mean += [cx, cy, s_k / sqrt(ar), s_k * sqrt(ar)]
    out = np.asarray(mean, dtype=np.float64).astype(F32).reshape(-1, 4)
    if cfg['clip']:
        out = np.clip(out, F32(0), F32(1))
    return out
