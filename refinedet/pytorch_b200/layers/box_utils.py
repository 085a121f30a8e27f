"""Drop-in for the reference's ``layers/box_utils.py`` backed by the sm_100a kernels.

Same names, argument order and results as the reference (file:line cited per function);
every function takes CUDA tensors and calls ``librefinedet_b200.so`` through ``_ffi``.
There is no CPU path.
"""
import torch

from .. import _ffi
from .._ffi import check, lib, ptr, require_cuda_f32, stream_ptr


def _out_like(t, shape=None, dtype=None):
    return torch.empty(t.shape if shape is None else shape, dtype=dtype or t.dtype, device=t.device)


def point_form(boxes):
    """layers/box_utils.py:5-14  (cx,cy,w,h) -> (x1,y1,x2,y2)."""
    b = require_cuda_f32(boxes, 'boxes')
    out = _out_like(b)
    with torch.cuda.device(b.device):
        check(lib().rd_point_form(ptr(b), ptr(out), b.shape[0], stream_ptr()), 'rd_point_form')
    return out


def center_size(boxes):
    """layers/box_utils.py:17-26  (x1,y1,x2,y2) -> (cx,cy,w,h)."""
    b = require_cuda_f32(boxes, 'boxes')
    out = _out_like(b)
    with torch.cuda.device(b.device):
        check(lib().rd_center_size(ptr(b), ptr(out), b.shape[0], stream_ptr()), 'rd_center_size')
    return out


def intersect(box_a, box_b):
    """layers/box_utils.py:29-47  [A,4] x [B,4] -> [A,B] intersection areas."""
    a, b = require_cuda_f32(box_a, 'box_a'), require_cuda_f32(box_b, 'box_b')
    out = torch.empty(a.shape[0], b.shape[0], dtype=torch.float32, device=a.device)
    with torch.cuda.device(a.device):
        check(lib().rd_intersect(ptr(a), ptr(b), ptr(out), a.shape[0], b.shape[0], stream_ptr()), 'rd_intersect')
    return out


def jaccard(box_a, box_b):
    """layers/box_utils.py:50-68  [A,4] x [B,4] -> [A,B] IoU."""
    a, b = require_cuda_f32(box_a, 'box_a'), require_cuda_f32(box_b, 'box_b')
    out = torch.empty(a.shape[0], b.shape[0], dtype=torch.float32, device=a.device)
    with torch.cuda.device(a.device):
        check(lib().rd_jaccard(ptr(a), ptr(b), ptr(out), a.shape[0], b.shape[0], stream_ptr()), 'rd_jaccard')
    return out


def encode(matched, priors, variances):
    """layers/box_utils.py:162-183."""
    m, p = require_cuda_f32(matched, 'matched'), require_cuda_f32(priors, 'priors')
    if m.shape != p.shape:
        raise ValueError('matched %s and priors %s must have the same shape' % (tuple(m.shape), tuple(p.shape)))
    out = _out_like(m)
    with torch.cuda.device(m.device):
        check(lib().rd_encode(ptr(m), ptr(p), float(variances[0]), float(variances[1]), ptr(out), m.shape[0],
                              stream_ptr()), 'rd_encode')
    return out


def decode(loc, priors, variances):
    """layers/box_utils.py:187-205."""
    l, p = require_cuda_f32(loc, 'loc'), require_cuda_f32(priors, 'priors')
    if l.shape != p.shape:
        raise ValueError('loc %s and priors %s must have the same shape' % (tuple(l.shape), tuple(p.shape)))
    out = _out_like(l)
    with torch.cuda.device(l.device):
        check(lib().rd_decode(ptr(l), ptr(p), float(variances[0]), float(variances[1]), ptr(out), l.shape[0],
                              stream_ptr()), 'rd_decode')
    return out


def log_sum_exp(x):
    """layers/box_utils.py:208-216 — stays on stock PyTorch (needs autograd)."""
    x_max = x.data.max()
    return torch.log(torch.sum(torch.exp(x - x_max), 1, keepdim=True)) + x_max


# ---------------------------------------------------------------------------------------------
# matching
# ---------------------------------------------------------------------------------------------
LABEL_ODM, LABEL_ARM_BINARY, LABEL_SSD_PLUS1 = 0, 1, 2


def pad_targets(targets, device):
    """``targets``: list of B tensors ``[G_i, 5]`` (x1,y1,x2,y2,label) -> padded
    ``truths[B,Gmax,4]``, ``labels[B,Gmax]``, ``gt_count[B]`` on ``device`` (the batched
    form of the per-image slicing at refinedet_multibox_loss.py:76-77)."""
    counts = [int(t.shape[0]) for t in targets]
    gmax = max(max(counts), 1)
    if gmax > _ffi.RD_MAX_GT:
        raise RuntimeError('more than %d ground-truth boxes in one image (%d)' % (_ffi.RD_MAX_GT, gmax))
    padded = torch.zeros(len(targets), gmax, 5, dtype=torch.float32, device=device)
    for i, t in enumerate(targets):
        if counts[i]:
            padded[i, :counts[i]] = t.detach().to(device=device, dtype=torch.float32)
    truths = padded[:, :, :4].contiguous()
    labels = padded[:, :, 4].contiguous()
    gt_count = torch.tensor(counts, dtype=torch.int32).to(device, non_blocking=True)
    return truths, labels, gt_count


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
    with torch.cuda.device(dev):
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
    with torch.cuda.device(loss_c.device):
        check(lib().rd_hnm_select(ptr(loss_c), ptr(pos), B, P, int(negpos_ratio), ptr(neg), ptr(num_pos),
                                  stream_ptr()), 'rd_hnm_select')
    return neg, num_pos


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
    with torch.cuda.device(dev):
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
