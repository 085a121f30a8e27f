"""Drop-in for the reference's ``layers/box_utils.py`` backed by the sm_100a kernels.

Same names, argument order and results as the reference (file:line cited per function);
every function takes CUDA tensors and calls ``librefinedet_b200.so`` through ``_ffi``.
There is no CPU path.
"""
import operator
import threading

import numpy as np
import torch

from .. import _ffi
from .._ffi import check, lib, on_device, ptr, require_cuda_f32, stream_ptr


def _out_like(t, shape=None, dtype=None):
    return torch.empty(t.shape if shape is None else shape, dtype=dtype or t.dtype, device=t.device)


def point_form(boxes):
    """layers/box_utils.py:5-14  (cx,cy,w,h) -> (x1,y1,x2,y2)."""
    b = require_cuda_f32(boxes, 'boxes')
    out = _out_like(b)
    with on_device(b.device):
        check(lib().rd_point_form(ptr(b), ptr(out), b.shape[0], stream_ptr()), 'rd_point_form')
    return out


def center_size(boxes):
    """layers/box_utils.py:17-26  (x1,y1,x2,y2) -> (cx,cy,w,h)."""
    b = require_cuda_f32(boxes, 'boxes')
    out = _out_like(b)
    with on_device(b.device):
        check(lib().rd_center_size(ptr(b), ptr(out), b.shape[0], stream_ptr()), 'rd_center_size')
    return out


def intersect(box_a, box_b):
    """layers/box_utils.py:29-47  [A,4] x [B,4] -> [A,B] intersection areas."""
    a, b = require_cuda_f32(box_a, 'box_a'), require_cuda_f32(box_b, 'box_b')
    out = torch.empty(a.shape[0], b.shape[0], dtype=torch.float32, device=a.device)
    with on_device(a.device):
        check(lib().rd_intersect(ptr(a), ptr(b), ptr(out), a.shape[0], b.shape[0], stream_ptr()), 'rd_intersect')
    return out


def jaccard(box_a, box_b):
    """layers/box_utils.py:50-68  [A,4] x [B,4] -> [A,B] IoU."""
    a, b = require_cuda_f32(box_a, 'box_a'), require_cuda_f32(box_b, 'box_b')
    out = torch.empty(a.shape[0], b.shape[0], dtype=torch.float32, device=a.device)
    with on_device(a.device):
        check(lib().rd_jaccard(ptr(a), ptr(b), ptr(out), a.shape[0], b.shape[0], stream_ptr()), 'rd_jaccard')
    return out


def encode(matched, priors, variances):
    """layers/box_utils.py:162-183."""
    m, p = require_cuda_f32(matched, 'matched'), require_cuda_f32(priors, 'priors')
    if m.shape != p.shape:
        raise ValueError('matched %s and priors %s must have the same shape' % (tuple(m.shape), tuple(p.shape)))
    out = _out_like(m)
    with on_device(m.device):
        check(lib().rd_encode(ptr(m), ptr(p), float(variances[0]), float(variances[1]), ptr(out), m.shape[0],
                              stream_ptr()), 'rd_encode')
    return out


def decode(loc, priors, variances):
    """layers/box_utils.py:187-205."""
    l, p = require_cuda_f32(loc, 'loc'), require_cuda_f32(priors, 'priors')
    if l.shape != p.shape:
        raise ValueError('loc %s and priors %s must have the same shape' % (tuple(l.shape), tuple(p.shape)))
    out = _out_like(l)
    with on_device(l.device):
        check(lib().rd_decode(ptr(l), ptr(p), float(variances[0]), float(variances[1]), ptr(out), l.shape[0],
                              stream_ptr()), 'rd_decode')
    return out


def conf_loss(conf, conf_t, arm_conf=None, theta=0.01):
    """One pass over ``conf[..., C]`` (refinedet_multibox_loss.py:96-101,113-114): returns
    ``(ce, lse, pos)`` shaped like ``conf_t`` — ``lse = log_sum_exp(conf)``, ``ce = lse - conf.gather(conf_t)``
    (the mining loss AND the cross-entropy term), ``pos = conf_t > 0`` gated by
    ``softmax(arm_conf)[..., 1] > theta`` when ``arm_conf`` (logits) is given.  No autograd."""
    conf = require_cuda_f32(conf, 'conf', align=8)
    C = conf.shape[-1]
    if conf_t.dtype != torch.int64 or not conf_t.is_cuda:
        raise TypeError('conf_t must be a CUDA int64 tensor')
    conf_t = conf_t.contiguous()
    rows = conf_t.numel()
    if conf.numel() != rows * C:
        raise ValueError('conf must hold conf_t.numel() rows of C values')
    if arm_conf is not None:
        arm_conf = require_cuda_f32(arm_conf, 'arm_conf', align=8)
        if arm_conf.numel() != rows * 2:
            raise ValueError('arm_conf must hold conf_t.numel() rows of 2 logits')
    ce = torch.empty(conf_t.shape, dtype=torch.float32, device=conf.device)
    lse = torch.empty_like(ce)
    pos = torch.empty(conf_t.shape, dtype=torch.bool, device=conf.device)
    with on_device(conf.device):
        check(lib().rd_conf_loss(ptr(conf), ptr(conf_t), ptr(arm_conf), float(theta), rows, C, ptr(ce), ptr(lse),
                                 ptr(pos), stream_ptr()), 'rd_conf_loss')
    return ce, lse, pos


def log_sum_exp(x):
    """layers/box_utils.py:208-216: ``log(sum(exp(x), 1, keepdim=True))`` of ``x[N,C]`` → ``[N,1]``, computed
    by ``rd_conf_loss`` (row max instead of the reference's global max: same value, no underflow).
    Values only — the loss module gets its gradients from ``rd_multibox_loss_backward``."""
    t = torch.zeros(x.shape[0], dtype=torch.int64, device=x.device)
    return conf_loss(x, t)[1].unsqueeze(1)


# ---------------------------------------------------------------------------------------------
# matching
# ---------------------------------------------------------------------------------------------
LABEL_ODM, LABEL_ARM_BINARY, LABEL_SSD_PLUS1 = 0, 1, 2
_is = operator.is_
_tls = threading.local()            # per-thread: the padded targets of the latest step (see _padded)


def clear_pad_cache():
    """Drop this thread's cached padded targets (they pin the step's target tensors until the next step)."""
    _tls.pad_cache = None



# Offsets of the ragged target list live in a small ring of PINNED host slots that the pad kernel reads directly: a
# copy from pageable memory would make the driver drain the stream first (CUDA API synchronisation rules), i.e. one
# hidden device synchronisation per training step.  A slot is reused only after the kernel that read it has completed.
_PIN_SLOTS = 8


def _offsets_slot(counts, device):
    """A pinned host slot holding the exclusive prefix sum of ``counts`` + the event that guards its reuse.  The pad
    kernel reads the offsets straight from the pinned slot (unified addressing: 33 ints over PCIe) — no H2D copy
    call at all; the caller records the event after launching the kernel.  The ring (and its cursor) is per thread
    and per (device, length)."""
    n = len(counts) + 1
    rings = getattr(_tls, 'pin_rings', None)
    if rings is None:
        rings = _tls.pin_rings = {}
    key = (device.index, n)
    ring = rings.get(key)
    if ring is None:
        ring = rings[key] = {'next': 0, 'slots': [(torch.zeros(n, dtype=torch.int32).pin_memory(), torch.cuda.Event())
                                                  for _ in range(_PIN_SLOTS)]}
    buf, ev = ring['slots'][ring['next'] % _PIN_SLOTS]
    ring['next'] += 1
    ev.synchronize()                                        # the kernel that last read the slot is done (normally long ago)
    np.cumsum(counts, out=buf.numpy()[1:])
    return buf, ev


def pad_targets(targets, device):
    """``targets``: list of B tensors ``[G_i, 5]`` (x1,y1,x2,y2,label) -> padded
    ``truths[B,Gmax,4]``, ``labels[B,Gmax]``, ``gt_count[B]`` on ``device`` (the batched
    form of the per-image slicing at refinedet_multibox_loss.py:76-77)."""
    return _padded(targets, device)[:3]


def _padded(targets, device):
    """:func:`pad_targets` plus the smallest per-image count (the host already knows it)."""
    # The ARM and the ODM criterion of a training step pad the same list (train_refinedet.py:252-253): the
    # latest result is kept.  The cache HOLDS the target tensors, so neither their ids nor their storage can
    # be recycled for another batch while the entry is alive; in-place edits show in ``_version``.
    device = torch.device(device)
    if device.index is None:
        device = torch.device(device.type, torch.cuda.current_device())
    c = getattr(_tls, 'pad_cache', None)
    if c is not None and c[0] == device and len(c[1]) == len(targets) and all(map(_is, targets, c[1])) \
            and [t._version for t in targets] == c[2]:
        return c[3]
    counts = [t.shape[0] for t in targets]
    gmax = max(max(counts), 1)
    if gmax > _ffi.RD_MAX_GT:
        raise RuntimeError('more than %d ground-truth boxes in one image (%d)' % (_ffi.RD_MAX_GT, gmax))
    B = len(targets)
    truths = torch.empty(B, gmax, 4, dtype=torch.float32, device=device)
    labels = torch.empty(B, gmax, dtype=torch.float32, device=device)
    gt_count = torch.empty(B, dtype=torch.int32, device=device)
    if sum(counts) == 0:
        truths.zero_(); labels.zero_(); gt_count.zero_()
    else:
        # one concatenation + one kernel instead of B slice assignments
        try:
            flat = torch.cat(targets)                     # same device everywhere, [G_i,5] rows: the common case
        except (RuntimeError, TypeError):
            flat = torch.cat([t.detach().reshape(-1, 5).to(device) for t in targets if t.shape[0]])
        flat = flat.detach().reshape(-1, 5)
        if flat.device != device or flat.dtype != torch.float32:
            flat = flat.to(device=device, dtype=torch.float32)
        flat = flat.contiguous()
        offsets, ev = _offsets_slot(counts, device)
        with on_device(device):
            check(lib().rd_pad_targets(ptr(flat), ptr(offsets), B, gmax, ptr(truths), ptr(labels), ptr(gt_count),
                                       stream_ptr()), 'rd_pad_targets')
            ev.record(torch.cuda.current_stream(device))
    out = (truths, labels, gt_count, min(counts))
    _tls.pad_cache = (device, list(targets), [t._version for t in targets], out)
    return out


def check_targets(targets, device=None):
    """The ground-truth sanity check of the reference's training loop (``train_refinedet.py:240-245``:
    a triple Python loop, one host sync per coordinate) as one device reduction and one sync:
    raises ``StopIteration`` when any box coordinate lies outside ``[0, 1]``.  Returns the padded
    ``(truths, labels, gt_count)`` so the loss does not pad again."""
    if device is None:
        device = targets[0].device
    truths, labels, gt_count = pad_targets(targets, device)
    G = truths.shape[1]
    valid = torch.arange(G, device=truths.device)[None, :] < gt_count[:, None]
    bad = ((truths < 0) | (truths > 1)).any(-1) & valid
    if bool(bad.any()):
        raise StopIteration
    return truths, labels, gt_count


def match_batch(threshold, truths, labels, gt_count, priors, variances, arm_loc=None,
                label_mode=LABEL_ODM, return_best=False):
    """Batched ``refine_match`` / ``match`` (box_utils.py:70-160) — one call for the whole
    batch instead of the per-image Python loop at refinedet_multibox_loss.py:75-86.

    ``truths[B,Gmax,4]``, ``labels[B,Gmax]``, ``gt_count[B]`` int32, ``priors[P,4]``,
    ``arm_loc[B,P,4]`` or None.  Returns ``loc_t[B,P,4]`` f32, ``conf_t[B,P]`` int64
    (plus ``best_truth_idx``/``best_truth_overlap`` when asked)."""
    truths = require_cuda_f32(truths, 'truths')
    labels = require_cuda_f32(labels, 'labels', align=4)
    priors = require_cuda_f32(priors, 'priors')
    dev = truths.device
    B, gmax = truths.shape[0], truths.shape[1]
    P = priors.shape[0]
    if arm_loc is not None:
        arm_loc = require_cuda_f32(arm_loc, 'arm_loc')
        if tuple(arm_loc.shape) != (B, P, 4):
            raise ValueError('arm_loc must be [B,P,4]')
    gt_count = gt_count.to(device=dev, dtype=torch.int32).contiguous()
    loc_t = torch.empty(B, P, 4, dtype=torch.float32, device=dev)
    conf_t = torch.empty(B, P, dtype=torch.int64, device=dev)
    bt_idx = torch.empty(B, P, dtype=torch.int32, device=dev)
    bt_ov = torch.empty(B, P, dtype=torch.float32, device=dev)
    L = lib()
    ws_bytes = int(L.rd_match_workspace_bytes(B, gmax))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    with on_device(dev):
        check(L.rd_refine_match(ptr(truths), ptr(labels), ptr(gt_count), ptr(priors), ptr(arm_loc), B, P, gmax,
                                float(threshold), float(variances[0]), float(variances[1]), int(label_mode),
                                ptr(ws), ws_bytes, ptr(loc_t), ptr(conf_t), ptr(bt_idx), ptr(bt_ov),
                                stream_ptr()), 'rd_refine_match')
    if return_best:
        return loc_t, conf_t, bt_idx, bt_ov
    return loc_t, conf_t


def _match_one(threshold, truths, priors, variances, labels, loc_t, conf_t, idx, arm_loc, label_mode):
    if truths.shape[0] == 0:
        raise IndexError('refine_match: no ground-truth boxes (the reference raises on max() over an empty dim)')
    dev = priors.device
    t = truths.detach().to(device=dev, dtype=torch.float32).reshape(1, -1, 4).contiguous()
    lab = labels.detach().to(device=dev, dtype=torch.float32).reshape(1, -1).contiguous()
    cnt = torch.tensor([t.shape[1]], dtype=torch.int32, device=dev)
    al = None if arm_loc is None else arm_loc.detach().reshape(1, -1, 4)
    l, c = match_batch(threshold, t, lab, cnt, priors, variances, al, label_mode)
    loc_t[idx] = l[0].to(loc_t.device)
    conf_t[idx] = c[0].to(conf_t.device)


def match(threshold, truths, priors, variances, labels, loc_t, conf_t, idx):
    """layers/box_utils.py:70-111 (SSD ``match``: conf = labels + 1)."""
    _match_one(threshold, truths, priors, variances, labels, loc_t, conf_t, idx, None, LABEL_SSD_PLUS1)


def refine_match(threshold, truths, priors, variances, labels, loc_t, conf_t, idx, arm_loc=None):
    """layers/box_utils.py:113-160.  ``labels`` may be float (1-based classes) or bool
    (``labels >= 0`` of the ARM criterion, refinedet_multibox_loss.py:78-79)."""
    if labels.dtype == torch.bool:
        # True -> 1, False -> 0: encode as +1 / -1 so that label_mode 1 (lab >= 0) restores it
        labels = labels.to(torch.float32) * 2 - 1
        mode = LABEL_ARM_BINARY
    else:
        mode = LABEL_ODM
    _match_one(threshold, truths, priors, variances, labels, loc_t, conf_t, idx, arm_loc, mode)


def hnm_select(loss_c, pos, negpos_ratio):
    """Hard-negative selection of refinedet_multibox_loss.py:117-123 without the two sorts.

    ``loss_c[B,P]`` f32 (entries at positives are treated as 0), ``pos[B,P]`` bool.
    Returns ``(neg[B,P] bool, num_pos[B] int32)``."""
    loss_c = require_cuda_f32(loss_c, 'loss_c', align=4)
    if pos.dtype != torch.bool or not pos.is_cuda:
        raise TypeError('pos must be a CUDA bool tensor')
    pos = pos.contiguous()
    B, P = loss_c.shape
    neg = torch.empty(B, P, dtype=torch.bool, device=loss_c.device)
    num_pos = torch.empty(B, dtype=torch.int32, device=loss_c.device)
    with on_device(loss_c.device):
        check(lib().rd_hnm_select(ptr(loss_c), ptr(pos), B, P, int(negpos_ratio), ptr(neg), ptr(num_pos),
                                  stream_ptr()), 'rd_hnm_select')
    return neg, num_pos


def multibox_loss_reduce(loc_data, loc_t, ce, pos, neg, num_pos):
    """``(loss_l, loss_c, N)`` 0-dim device tensors: SmoothL1 over ``pos`` and ``ce`` over ``pos | neg``,
    both divided by ``N = sum(num_pos)`` (refinedet_multibox_loss.py:105-110,126-138); zeros when ``N < 1``."""
    loc_data = require_cuda_f32(loc_data, 'loc_data')
    loc_t = require_cuda_f32(loc_t, 'loc_t')
    B, P = ce.shape
    dev = ce.device
    L = lib()
    ws = torch.empty(int(L.rd_multibox_loss_workspace_bytes(B)), dtype=torch.uint8, device=dev)
    out = [torch.empty((), dtype=torch.float32, device=dev) for _ in range(3)]
    with on_device(dev):
        check(L.rd_multibox_loss_reduce(ptr(loc_data), ptr(loc_t), ptr(ce), ptr(pos), ptr(neg), ptr(num_pos), B, P,
                                        ptr(ws), ws.numel(), ptr(out[0]), ptr(out[1]), ptr(out[2]), stream_ptr()),
              'rd_multibox_loss_reduce')
    return out[0], out[1], out[2]


def multibox_loss_backward(loc_data, loc_t, conf_data, conf_t, lse, pos, neg, g_l, g_c, n_dev,
                           need_loc=True, need_conf=True):
    """Gradients of ``(loss_l, loss_c)`` w.r.t. ``(loc_data, conf_data)``; ``g_l`` / ``g_c`` are 0-dim device
    tensors (or None)."""
    C = conf_data.shape[-1]
    rows = conf_t.numel()
    grad_loc = torch.empty_like(loc_data) if need_loc else None
    grad_conf = torch.empty_like(conf_data) if need_conf else None
    with on_device(conf_data.device):
        check(lib().rd_multibox_loss_backward(ptr(loc_data), ptr(loc_t), ptr(conf_data), ptr(conf_t), ptr(lse),
                                              ptr(pos), ptr(neg), ptr(g_l), ptr(g_c), ptr(n_dev), rows, C,
                                              ptr(grad_loc), ptr(grad_conf), stream_ptr()),
              'rd_multibox_loss_backward')
    return grad_loc, grad_conf


# ---------------------------------------------------------------------------------------------
# threshold + top-k select
# ---------------------------------------------------------------------------------------------
def select_topk(scores, conf_thresh, top_k, first_class=1):
    """Per (image, class) candidate lists of the detect stage as the reference builds them before NMS:
    ``scores[b,:,c] > conf_thresh`` (eval_refinedet_coco.py:214, detection_refinedet.py:98), the ``top_k``
    highest in score-descending order (eval :222, box_utils.py:242-244), lower anchor first on ties.

    ``scores[B,P,C]`` (e.g. the second output of ``Detect_RefineDet.forward``).  Returns device tensors
    ``(idx[B,C,top_k] int32, sc[B,C,top_k] f32, counts[B,C] int32)``; only the first ``counts[b,c]``
    entries of a slot are written; classes below ``first_class`` (background) get count 0.  No host sync."""
    s = require_cuda_f32(scores, 'scores', align=4)
    if s.dim() != 3:
        raise ValueError('scores must be [B,P,C], got %s' % (tuple(s.shape),))
    B, P, C = (int(v) for v in s.shape)
    top_k = int(top_k)
    if top_k <= 0:
        raise ValueError('top_k must be positive')
    if min(top_k, P) > 4 * _ffi.RD_MAX_NMS_BOXES:
        raise RuntimeError('select_topk: min(top_k, P) = %d exceeds the supported %d'
                           % (min(top_k, P), 4 * _ffi.RD_MAX_NMS_BOXES))
    dev = s.device
    idx = torch.empty(B, C, top_k, dtype=torch.int32, device=dev)
    sc = torch.empty(B, C, top_k, dtype=torch.float32, device=dev)
    counts = torch.empty(B, C, dtype=torch.int32, device=dev)
    if B == 0 or P == 0 or C == 0:
        return idx, sc, counts.zero_()
    with on_device(dev):
        check(lib().rd_select_topk(ptr(s), B, P, C, float(conf_thresh), top_k, int(first_class), ptr(idx), ptr(sc),
                                   ptr(counts), stream_ptr()), 'rd_select_topk')
    return idx, sc, counts


# ---------------------------------------------------------------------------------------------
# NMS
# ---------------------------------------------------------------------------------------------
def nms_device(boxes, scores, overlap, top_k, flags=_ffi.RD_NMS_NORMALISED):
    """Device-side NMS without a host sync: returns ``(keep[int64, n], count[int32, 1])``
    tensors; ``keep[:count]`` are the kept original indices, score-descending."""
    boxes = require_cuda_f32(boxes, 'boxes')
    scores = require_cuda_f32(scores, 'scores', align=4)
    n = int(scores.shape[0])
    dev = boxes.device
    keep = torch.zeros(n, dtype=torch.int64, device=dev)
    count = torch.zeros(1, dtype=torch.int32, device=dev)
    if n == 0:
        return keep, count
    if min(int(top_k), n) > _ffi.RD_MAX_NMS_BOXES:
        raise RuntimeError('nms: min(top_k, n) = %d exceeds the supported %d boxes per problem'
                           % (min(int(top_k), n), _ffi.RD_MAX_NMS_BOXES))
    L = lib()
    ws_bytes = int(L.rd_nms_workspace_bytes(n))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    with on_device(dev):
        check(L.rd_nms(ptr(boxes), ptr(scores), n, float(overlap), int(top_k), int(flags), ptr(ws), ws_bytes,
                       ptr(keep), ptr(count), stream_ptr()), 'rd_nms')
    return keep, count


def nms(boxes, scores, overlap=0.5, top_k=200):
    """layers/box_utils.py:222-286.  Returns ``(keep, count)`` like the reference —
    ``keep`` int64 of length n, zero padded, ``count`` a Python int — and, like the
    reference (:235-236), the bare ``keep`` tensor when ``boxes`` is empty."""
    if boxes.numel() == 0:
        return scores.new_zeros(scores.size(0), dtype=torch.long)
    keep, count = nms_device(boxes, scores, overlap, top_k, _ffi.RD_NMS_NORMALISED)
    return keep, int(count.item())
