"""Drop-in for ``layers/modules/refinedet_multibox_loss.py`` (reference :10-139).

The per-image ``refine_match`` Python loop (:75-86) becomes one batched kernel pair, the double
sort of the hard-negative mining (:119-123) one radix-select kernel, and the loss tail
(:96-138: ARM-theta gate, ``log_sum_exp - gather``, SmoothL1, cross-entropy, ``/ N``) three
kernels forward and one backward behind a ``torch.autograd.Function`` — ``conf_data`` is read once
forward and only on the ``pos | neg`` rows backward.
"""
import torch
import torch.nn as nn

from ..._ffi import require_cuda_f32
from ..box_utils import (LABEL_ARM_BINARY, LABEL_ODM, conf_loss, hnm_select, match_batch,
                         multibox_loss_backward, multibox_loss_reduce, _padded)

# data/config.py:57 (``coco['variance']``, read at reference :46)
_VARIANCE = [0.1, 0.2]


class RefineDetMultiBoxLoss(nn.Module):
    """RefineDet weighted loss: SmoothL1 localisation + cross-entropy confidence with 3:1 hard
    negative mining, for the ARM (``use_ARM=False``, 2-class) or ODM (``use_ARM=True``)
    branch.  Constructor arguments are the reference's (:33-48).

    ``sync_free`` (extension, default off): the reference reads ``N = sum(num_pos)`` on the host to return
    ``(zeros(1), zeros(1))`` when no positive survives (:135-136) — a stream drain in the middle of every
    criterion call.  With ``sync_free=True`` nothing is read back: the losses are always the 0-dim device
    tensors, and for ``N < 1`` the kernels deliver exactly zero losses and zero gradients — the same parameter
    update as the reference's no-grad zeros, so ``train_refinedet.py:252-261`` runs unchanged while the host
    queues the next criterion and the backward pass behind the running kernels."""

    def __init__(self, num_classes, overlap_thresh, prior_for_matching, bkg_label, neg_mining,
                 neg_pos, neg_overlap, encode_target, use_gpu=True, theta=0.01, use_ARM=False, sync_free=False):
        super(RefineDetMultiBoxLoss, self).__init__()
        self.use_gpu = use_gpu
        self.num_classes = num_classes
        self.threshold = overlap_thresh
        self.background_label = bkg_label
        self.encode_target = encode_target
        self.use_prior_for_matching = prior_for_matching
        self.do_neg_mining = neg_mining
        self.negpos_ratio = neg_pos
        self.neg_overlap = neg_overlap
        self.variance = _VARIANCE
        self.theta = theta
        self.use_ARM = use_ARM
        self.sync_free = bool(sync_free)

    def match_targets(self, predictions, targets):
        """Targets of reference :62-90 for the whole batch: ``(loc_t[B,P,4], conf_t[B,P])``."""
        arm_loc_data, arm_conf_data, odm_loc_data, odm_conf_data, priors = predictions
        loc_data = odm_loc_data if self.use_ARM else arm_loc_data
        if not loc_data.is_cuda:
            raise RuntimeError('RefineDetMultiBoxLoss: predictions must be CUDA tensors '
                               '(refinedet.pytorch_b200 has no CPU fallback)')
        truths, labels, gt_count, min_count = _padded(targets, loc_data.device)
        if min_count == 0:
            raise IndexError('RefineDetMultiBoxLoss: an image has no ground-truth boxes '
                             '(the reference raises in refine_match, box_utils.py:139)')
        priors = priors[:loc_data.size(1), :]                       # :68 (DataParallel gather)
        if self.num_classes == 2 and not self.use_ARM:
            mode = LABEL_ARM_BINARY                                 # labels = labels >= 0, :78-79
        else:
            mode = LABEL_ODM
        arm = arm_loc_data.detach() if self.use_ARM else None       # :80-85
        return match_batch(self.threshold, truths, labels, gt_count, priors.detach(), self.variance, arm, mode)

    def forward(self, predictions, targets):
        """reference :50-139.  ``predictions`` = (arm_loc, arm_conf, odm_loc, odm_conf, priors),
        ``targets`` = list of ``[G_i,5]`` tensors.  Returns ``(loss_l, loss_c)``."""
        arm_loc_data, arm_conf_data, odm_loc_data, odm_conf_data, priors = predictions
        if self.use_ARM:
            loc_data, conf_data = odm_loc_data, odm_conf_data
        else:
            loc_data, conf_data = arm_loc_data, arm_conf_data
        loc_t, conf_t = self.match_targets(predictions, targets)
        arm_gate = arm_conf_data.detach() if self.use_ARM else None           # :96-101 (softmax inside the kernel)
        loss_l, loss_c, N, pos, neg = _MultiBoxLossTail.apply(loc_data, conf_data, arm_gate, loc_t, conf_t,
                                                              float(self.theta), int(self.negpos_ratio))
        self.last_masks = (pos, neg)        # of this criterion's latest forward, for inspection / tests
        if not self.sync_free and float(N) < 1:                     # :135-136 (the reference syncs here too)
            return torch.zeros(1), torch.zeros(1)
        return loss_l, loss_c


class _MultiBoxLossTail(torch.autograd.Function):
    """refinedet_multibox_loss.py:96-138 on the device: forward = rd_conf_loss + rd_hnm_select +
    rd_multibox_loss_reduce, backward = rd_multibox_loss_backward."""

    @staticmethod
    def forward(ctx, loc_data, conf_data, arm_conf, loc_t, conf_t, theta, negpos_ratio):
        loc_c = require_cuda_f32(loc_data, 'loc_data')
        conf_c = require_cuda_f32(conf_data, 'conf_data', align=8)
        B, P = conf_t.shape
        ce, lse, pos = conf_loss(conf_c, conf_t, arm_conf, theta)              # :96-101, :113-114
        neg, num_pos = hnm_select(ce, pos, negpos_ratio)                       # :117-123
        loss_l, loss_c, N = multibox_loss_reduce(loc_c, loc_t, ce, pos, neg, num_pos)   # :105-110, :126-138
        ctx.save_for_backward(loc_c, conf_c, loc_t, conf_t, lse, pos, neg, N)
        ctx.mark_non_differentiable(N, pos, neg)
        return loss_l, loss_c, N, pos, neg

    @staticmethod
    def backward(ctx, g_l, g_c, _g_n, _g_pos, _g_neg):
        loc_c, conf_c, loc_t, conf_t, lse, pos, neg, N = ctx.saved_tensors
        need_loc, need_conf = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        g_l = g_l.contiguous().float() if g_l is not None else None
        g_c = g_c.contiguous().float() if g_c is not None else None
        grad_loc, grad_conf = multibox_loss_backward(loc_c, loc_t, conf_c, conf_t, lse, pos, neg, g_l, g_c, N,
                                                     need_loc, need_conf)
        return grad_loc, grad_conf, None, None, None, None, None
